"""Build-container only: the oracle against the reference itself, run here.  Skipped where /root/reference
or oracle/_ref is absent (the GPU box); tests/golden carries the same pin there."""
import os
import subprocess

import numpy as np
import pytest

from conftest import REF_BUILD, REFERENCE, SEED_SETS

pytestmark = pytest.mark.needs_reference
have_ref = os.path.isdir(REFERENCE) and os.path.exists(os.path.join(REF_BUILD, "bin", "base", "CLSuperPathTracer"))


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="needs /root/reference")
def test_scene_files_regenerate_byte_identical():
    import write_scenes
    assert write_scenes.verify_against_reference(REFERENCE)


@pytest.mark.skipif(not have_ref, reason="needs oracle/_ref (make -C oracle ref)")
@pytest.mark.parametrize("variant,size", [("base", (256, 384)), ("lmem", (256, 384)), ("nodof", (256, 384)), ("grid", (256, 384)),
                                          ("bidir", (320, 384))])
def test_result_ppm_identical(oracle_sep, scene_dirs, tmp_path, variant, size):
    """Unmodified reference host + kernel (CPU, refrt) vs oracle_render: identical result.ppm bytes."""
    w, h = size
    d = scene_dirs[variant]
    env = dict(os.environ, PT_SEEDS=",".join(str(s) for s in SEED_SETS[1]))
    exe = "CLSuperBidirectionalPathTracer" if variant == "bidir" else "CLSuperPathTracer"
    subprocess.run([os.path.join(REF_BUILD, "bin", variant, exe), str(w), str(h)], cwd=d, env=env, check=True,
                   capture_output=True)
    raw = open(os.path.join(d, "result.ppm"), "rb").read()
    k = raw.index(b"ENDHDR\n") + 7
    ref_img = np.frombuffer(raw[k:], np.uint8).reshape(h, w, 4)
    out = oracle_sep.render(variant, w, h, SEED_SETS[1], oracle_sep.load_scene_dir(d, variant), want_rng=False, want_accum=False)
    assert np.array_equal(out["image"], ref_img)
    oracle_sep.save_pam(os.path.join(str(tmp_path), "o.ppm"), out["image"])
    assert open(os.path.join(str(tmp_path), "o.ppm"), "rb").read() == raw
