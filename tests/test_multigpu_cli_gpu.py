"""Single-process multi-GPU path of the drop-in executables (PT_GPUS=n): row stripes + one NCCL reduce of the
accumulation buffer must give a result.ppm byte-identical to the single-GPU run.  Needs >= 2 GPUs (skipped on
the 1-GPU test box; exercised with `gpurun --gpus 2`)."""
import os
import re
import subprocess

import pytest

from conftest import ROOT
from opencl_montecarlo_path_tracing_b200 import _lib

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "opencl_montecarlo_path_tracing_b200", "bin")


@pytest.mark.parametrize("variant,dirname", [("nodof", "CLSuperPathTracer_lmem_NoDoF"), ("grid", "CLSuperPathTracer_trianglegrid"),
                                             ("base", "CLSuperPathTracer"), ("bidir", "CLSuperBidirectionalPathTracer")])
def test_pt_gpus_is_bit_identical(scene_dirs, variant, dirname):
    ngpu = _lib.cuda_lib().pt_device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    exe = os.path.join(BIN, dirname, "CLSuperBidirectionalPathTracer" if variant == "bidir" else "CLSuperPathTracer")
    d = scene_dirs[variant]
    imgs = {}
    for n in (1, min(ngpu, 8)):
        env = dict(os.environ, PT_SEEDS="1,2,3,4", PT_GPUS=str(n), PT_STATS="1")
        p = subprocess.run([exe, "640", "360"], cwd=d, env=env, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stdout + p.stderr
        imgs[n] = open(os.path.join(d, "result.ppm"), "rb").read()
        print(variant, n, re.search(r"rendering : .*", p.stdout).group(0))
    a, b = imgs.values()
    assert a == b


def test_pt_gpus_sample_sharding_matches_oracle_blocks(scene_dirs, oracle_fma):
    """PT_GPUS=2 PT_SHARD=samples: GPU i renders sample block i of the whole image; the NCCL sum of two buffers is exact,
    so result.ppm equals the tone-mapped sum of the oracle's two blocks."""
    import numpy as np
    ngpu = _lib.cuda_lib().pt_device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    d = scene_dirs["lmem"]
    exe = os.path.join(BIN, "CLSuperPathTracer_lmem", "CLSuperPathTracer")
    env = dict(os.environ, PT_SEEDS="1,2,3,4", PT_GPUS="2", PT_SHARD="samples")
    p = subprocess.run([exe, "256", "160"], cwd=d, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    raw = open(os.path.join(d, "result.ppm"), "rb").read()
    img = np.frombuffer(raw[raw.index(b"ENDHDR\n") + 7:], np.uint8).reshape(160, 256, 4)
    sc = oracle_fma.load_scene_dir(d, "lmem")
    parts = [oracle_fma.render("lmem", 256, 160, (1, 2, 3, 4), sc, sample_block=b, sample_blocks=2, want_rng=False)["accum"] for b in range(2)]
    total = parts[0] + parts[1]
    expect = np.clip(np.trunc(total), 0, 255).astype(np.uint8)
    assert np.array_equal(img, expect)
