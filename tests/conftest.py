import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scenes"))

REFERENCE = "/root/reference"          # only exists in the build container
REF_BUILD = os.path.join(ROOT, "oracle", "_ref")

SEED_SETS = [(1, 2, 3, 4), (123456789, 42, 7, 99999)]

# The parity tests compare the WORK COUNTERS with the oracle's as well, so by default they run the kernels in
# PT_DEAD_RAYS_TRACE mode (every ray of the reference is traced; include/ptcuda.h).  The library default — shadow rays of
# triangle-material samples elided, same image / accumulation / RNG bits — is what tests/test_dead_rays_gpu.py pins, with
# an explicit dead_rays="elide" and through the drop-in executables with this variable removed.
os.environ["PT_DEAD_RAYS"] = "trace"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference and oracle/_ref (build container only)")


@pytest.fixture(scope="session")
def scene_dirs(tmp_path_factory):
    """The five reference scene directories, regenerated from scenes/scene_data.py."""
    import write_scenes
    base = tmp_path_factory.mktemp("scenes")
    out = {}
    for v in ("base", "lmem", "nodof", "grid", "bidir"):
        d = str(base / v)
        write_scenes.write_variant(v, d)
        out[v] = d
    d = str(base / "torus")
    write_scenes.write_variant("base", d, mesh="torus")
    out["torus"] = d
    return out


@pytest.fixture(scope="session")
def oracle_fma():
    from oracle.pyoracle import OracleLib, cpu_has_fma
    if not cpu_has_fma():
        pytest.skip("host CPU lacks FMA")
    return OracleLib(contract=1)


@pytest.fixture(scope="session")
def oracle_sep():
    from oracle.pyoracle import OracleLib
    return OracleLib(contract=0)


def have_gpu():
    try:
        import opencl_montecarlo_path_tracing_b200._lib as L
        return L.cuda_lib().pt_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def renderer():
    import opencl_montecarlo_path_tracing_b200 as pt
    r = pt.Renderer(device=0)
    yield r
    r.close()
