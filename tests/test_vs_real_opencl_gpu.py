"""The UNMODIFIED reference (host .c + kernel .ocl) executed by NVIDIA's real OpenCL runtime on the B200 itself,
compared with this repo's CUDA path on the same seeds.  The binaries come from `make -C oracle ref`
(oracle/_ref/ocl/, kernel text embedded at build time; only buildable where /root/reference exists) and need
the driver's OpenCL ICD (OCL_ICD_FILENAMES=libnvidia-opencl.so.1).  Skipped when either is missing.

A real OpenCL compiler contracts, and divides / takes square roots with ~2 ulp error, so this comparison is
tolerance-based: the RNG-driven image must agree except where the reference's own rounding
noise decides visibility (squares self-shadow depending on the last bit of the hit point's z)."""
import os
import re
import subprocess

import numpy as np
import pytest

import opencl_montecarlo_path_tracing_b200 as pt
from conftest import ROOT, SEED_SETS

pytestmark = pytest.mark.gpu


def _ocl_exe(variant):
    return os.path.join(ROOT, "oracle", "_ref", "ocl", variant, "CLSuperPathTracer")


def _run_reference_opencl(variant, d, w, h, seeds):
    env = dict(os.environ, OCL_ICD_FILENAMES="libnvidia-opencl.so.1", PT_SEEDS=",".join(map(str, seeds)))
    p = subprocess.run([_ocl_exe(variant), str(w), str(h)], cwd=d, env=env, capture_output=True, text=True, timeout=600)
    if p.returncode != 0 or "selected device" not in p.stdout:
        pytest.skip("NVIDIA OpenCL ICD not usable here: " + (p.stdout + p.stderr)[-300:])
    raw = open(os.path.join(d, "result.ppm"), "rb").read()
    k = raw.index(b"ENDHDR\n") + 7
    ms = sum(float(x) for x in re.findall(r"(?:rendering|reduce img samples) : .*? in ([0-9.eE+-]+)ms", p.stdout))
    return np.frombuffer(raw[k:], np.uint8).reshape(h, w, 4), ms


@pytest.mark.parametrize("variant", ["base", "lmem", "nodof", "grid"])
def test_image_matches_reference_run_by_nvidia_opencl(renderer, scene_dirs, variant):
    if not os.path.exists(_ocl_exe(variant)):
        pytest.skip("oracle/_ref/ocl not built (needs /root/reference)")
    W = H = 512
    d = scene_dirs[variant]
    ref_img, ref_ms = _run_reference_opencl(variant, d, W, H, SEED_SETS[0])
    scene = pt.load_scene_dir(d, variant)
    renderer.set_scene(scene)
    if variant == "grid":
        renderer.build_grid(pt.grid_dims(scene))
    res = renderer.render(variant, W, H, SEED_SETS[0])
    diff = np.abs(res.image.astype(np.int32) - ref_img.astype(np.int32))[..., :3]
    within1 = float((diff <= 1).mean())
    rmse = float(np.sqrt((diff.astype(np.float64) ** 2).mean()))
    flipped = float((diff.max(axis=2) > 1).mean())
    print("%s: OpenCL %.3f ms, CUDA %.3f ms, within 1 LSB %.5f, RMSE %.4f LSB, pixels off by >1: %.5f" % (
        variant, ref_ms, res.ms, within1, rmse, flipped))
    assert (res.image[..., 3] == 255).all() and (ref_img[..., 3] == 255).all()
    assert within1 >= 0.995 and rmse <= 1.0 and flipped <= 0.005
    if variant == "base":          # no self-shadowing squares are lit in the base scene: essentially identical
        assert float((diff.max(axis=2) == 0).mean()) >= 0.999
