"""CLSuperBidirectionalPathTracer on the GPU, through the C ABI, against
  * the CPU oracle on the same seeds (bit-exact: VPL buffer, image, accumulation buffer, RNG state, counters), and
  * tests/golden/golden_bidir.npz, which was produced by the unmodified reference itself.
"""
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

import opencl_montecarlo_path_tracing_b200 as pt
from conftest import ROOT, SEED_SETS

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_bidir.npz")
BIN = os.path.join(ROOT, "opencl_montecarlo_path_tracing_b200", "bin", "CLSuperBidirectionalPathTracer", "CLSuperBidirectionalPathTracer")


def same_floats(a, b):
    """Bit equality, except that any NaN equals any NaN (x86 and the GPU produce different NaN payloads for 0/0)."""
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return a.shape == b.shape and np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


@pytest.fixture(scope="module")
def bidir(renderer, scene_dirs):
    scene = pt.load_scene_dir(scene_dirs["bidir"], "bidir")
    renderer.set_scene(scene)
    return renderer


@pytest.mark.parametrize("arith", ["fma", "separate"])
def test_light_tracer_bit_exact(bidir, scene_dirs, oracle_fma, oracle_sep, arith):
    o = oracle_fma if arith == "fma" else oracle_sep
    osc = o.load_scene_dir(scene_dirs["bidir"], "bidir")
    g = np.load(G)
    for si, seeds in enumerate(SEED_SETS):
        for n in (512, 700, 96, 1, 33):
            bidir.light_tracer(seeds, n, arith=arith)
            got = bidir.read_vpls()
            assert got.shape == (2 * n, 4)
            assert same_floats(got, o.light_tracer(seeds, osc, n)), (arith, si, n)
            if arith == "separate" and n in (512, 700, 96):           # the reference's own buffer
                assert same_floats(got, g["vpl_s%d_n%d" % (si, n)].view(np.float32)), ("golden", si, n)


@pytest.mark.parametrize("arith", ["fma", "separate"])
@pytest.mark.parametrize("mem", ["const", "smem"])
def test_pathtracer_bit_exact_windows(bidir, scene_dirs, oracle_fma, oracle_sep, arith, mem):
    o = oracle_fma if arith == "fma" else oracle_sep
    osc = o.load_scene_dir(scene_dirs["bidir"], "bidir")
    W = H = 512
    seeds = SEED_SETS[0]
    bidir.light_tracer(seeds, 512, arith=arith)
    vpls = o.light_tracer(seeds, osc, 512)
    for rows in ((112, 136), (196, 228), (340, 364), (500, 512)):
        res = bidir.render("bidir", W, H, seeds, rows=rows, arith=arith, scene_mem=mem, want_accum=True, want_rng=True)
        ref = o.render("bidir", W, H, seeds, osc, rows=rows, vpls=vpls)
        r0, r1 = rows
        assert np.array_equal(res.rng_state.reshape(H, W, 4)[r0:r1], ref["rng_state"].reshape(H, W, 4)[r0:r1]), rows
        assert same_floats(res.accum[r0:r1], ref["accum"][r0:r1]), rows
        assert np.array_equal(res.image[r0:r1], ref["image"][r0:r1]), rows
        for k in ("samples", "rays", "shadow_rays", "tri_tests", "prim_tests"):
            assert res.counters[k] == ref["counters"][k], (k, rows)
        assert res.counters["vpl_evals"] == ref["counters"]["shadow_rays"] // 2 * 1024


def test_reference_goldens_separate_policy(bidir):
    """CUDA (arith=separate) against bytes the unmodified reference wrote: selected rows of two 512x512 frames, a
    frame with N_VLP=700, and the whole degenerate N_VLP=96 frame whose VPLs carry inf/NaN intensities."""
    g = np.load(G)
    rows = [int(r) for r in g["rows"]]
    for si, seeds in enumerate(SEED_SETS):
        bidir.light_tracer(seeds, 512, arith="separate")
        for ri, r in enumerate(rows):
            res = bidir.render("bidir", 512, 512, seeds, rows=(r, r + 1), arith="separate")
            assert np.array_equal(res.image[r], g["img_s%d_rows" % si][ri]), (si, r)
    res = bidir.render("bidir", 512, 512, SEED_SETS[1], arith="separate")
    assert hashlib.sha256(res.image.tobytes()).digest() == g["img_s1_sha256"].tobytes()
    for n in (700, 96):
        w, h = (int(x) for x in g["img_n%d_size" % n])
        bidir.light_tracer(SEED_SETS[0], n, arith="separate")
        res = bidir.render("bidir", w, h, SEED_SETS[0], arith="separate")
        assert hashlib.sha256(res.image.tobytes()).digest() == g["img_n%d_sha256" % n].tobytes(), n
    assert np.array_equal(res.image, g["img_n96"])


def test_set_vpls_and_compaction(bidir, scene_dirs, oracle_fma):
    """Caller-supplied VPL buffers: zero-intensity entries anywhere (skipped like the reference's `continue`),
    -0 intensity, NaN/inf intensity, an empty buffer, and more entries than one compaction pass holds."""
    osc = oracle_fma.load_scene_dir(scene_dirs["bidir"], "bidir")
    rng = np.random.default_rng(7)
    W, H, rows = 512, 512, (352, 360)
    base = oracle_fma.light_tracer(SEED_SETS[0], osc, 512)
    cases = {"golden-like": base}
    dense = np.zeros((1500, 4), np.float32)
    dense[:, 0] = rng.uniform(0, 18, 1500); dense[:, 1] = rng.uniform(-3, 6, 1500); dense[:, 2] = rng.uniform(0, 13, 1500)
    dense[:, 3] = np.where(rng.uniform(size=1500) < 0.4, 0.0, rng.uniform(0.01, 0.3, 1500))
    dense[7, 3] = -0.0
    cases["dense"] = dense
    weird = base.copy()
    weird[3, 3] = np.inf; weird[900, 3] = np.nan; weird[5] = (1e30, 1e30, 1e30, 1.0)
    cases["inf-nan-far"] = weird
    cases["empty"] = np.zeros((0, 4), np.float32)
    cases["all-zero"] = np.zeros((64, 4), np.float32)
    for name, v in cases.items():
        bidir.set_vpls(v)
        assert same_floats(bidir.read_vpls(), v.reshape(-1, 4))
        res = bidir.render("bidir", W, H, SEED_SETS[0], rows=rows, want_accum=True)
        ref = oracle_fma.render("bidir", W, H, SEED_SETS[0], osc, rows=rows, vpls=v if len(v) else np.zeros((0, 4), np.float32))
        assert same_floats(res.accum[rows[0]:rows[1]], ref["accum"][rows[0]:rows[1]]), name
        assert np.array_equal(res.image[rows[0]:rows[1]], ref["image"][rows[0]:rows[1]]), name


def test_vpl_exactly_at_a_hit_point(bidir, scene_dirs, oracle_fma):
    """dist == 0 makes the reference divide 0/0: the flat-normal shortcut must fall back to the full expression."""
    osc = oracle_fma.load_scene_dir(scene_dirs["bidir"], "bidir")
    W, H, rows = 512, 512, (400, 402)
    probe = oracle_fma.render("bidir", W, H, SEED_SETS[0], osc, rows=rows, vpls=np.zeros((0, 4), np.float32))
    assert probe["counters"]["shadow_rays"] > 0
    # floor points the first samples of this window hit are not known in closed form; instead place VPLs ON the floor
    # plane (z = +-0 and denormal offsets), which exercises dv.z == +-0 and tiny dist for flat normals
    v = np.array([[8, 1, 0.0, 5], [9, 2, -0.0, 5], [10, 3, 1e-42, 5], [7, 0, -1e-42, 5], [17, 16, 8, 50]], np.float32)
    bidir.set_vpls(v)
    res = bidir.render("bidir", W, H, SEED_SETS[0], rows=rows, want_accum=True)
    ref = oracle_fma.render("bidir", W, H, SEED_SETS[0], osc, rows=rows, vpls=v)
    assert same_floats(res.accum[rows[0]:rows[1]], ref["accum"][rows[0]:rows[1]])


@pytest.mark.parametrize("nlights", [0, 1, 5])
def test_other_light_counts(renderer, scene_dirs, oracle_fma, nlights):
    scene = pt.load_scene_dir(scene_dirs["bidir"], "bidir")
    lights = np.array([[10, 4, 10, 200], [15, 2, 7, 150], [3, -2, 14, 90], [8, 5, 3, 60], [14, 0, 13.5, 300]], np.float32)[:nlights]
    scene.lights = lights.reshape(nlights, 4)
    renderer.set_scene(scene)
    osc = oracle_fma.load_scene_dir(scene_dirs["bidir"], "bidir")
    osc["lights"] = scene.lights
    renderer.light_tracer(SEED_SETS[1], 300)
    got = renderer.read_vpls()
    assert got.shape == (300 * nlights, 4)
    assert same_floats(got, oracle_fma.light_tracer(SEED_SETS[1], osc, 300))
    rows = (344, 352)
    res = renderer.render("bidir", 512, 512, SEED_SETS[1], rows=rows, want_accum=True)
    ref = oracle_fma.render("bidir", 512, 512, SEED_SETS[1], osc, rows=rows, n_vlp=300)
    assert same_floats(res.accum[rows[0]:rows[1]], ref["accum"][rows[0]:rows[1]])
    renderer.set_scene(pt.load_scene_dir(scene_dirs["bidir"], "bidir"))


def test_fast_math_is_exact(renderer):
    """The branch-free division / square root of the gather against the library's IEEE functions."""
    tested = 0
    for seed in (1, 2, 3, 4):
        out = renderer.selftest_fastmath(1 << 30, seed)
        assert out["div_mismatches"] == 0 and out["sqrt_mismatches"] == 0, out
        tested += out["pairs_tested"]
    assert tested > 3 << 30


def test_errors(scene_dirs):
    with pt.Renderer(0) as r:
        r.set_scene(pt.load_scene_dir(scene_dirs["bidir"], "bidir"))
        with pytest.raises(pt.PtError, match="pt_launch_lighttracer"):
            r.render("bidir", 64, 64, SEED_SETS[0])
        with pytest.raises(pt.PtError):
            r.read_vpls()
        with pytest.raises(pt.PtError):
            r.light_tracer(SEED_SETS[0], 0)
        r.light_tracer(SEED_SETS[0], 64)
        with pytest.raises(pt.PtError, match="megakernel"):
            r.render("bidir", 64, 64, SEED_SETS[0], kernel="wavefront")


def test_cli_dropin(scene_dirs, oracle_fma):
    d = scene_dirs["bidir"]
    env = dict(os.environ, PT_SEEDS="1,2,3,4")
    p = subprocess.run([BIN, "384", "320", "700"], cwd=d, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    out = p.stdout
    pos = 0
    for token in ["Usage:", "[N_VLP_per_light]", "number of platforms:", "selected device", "bidirectionalpathtracer.ocl", "Seeds: 1, 2, 3, 4",
                  "Processing image 384x320", "Cam values:", "Number of triangles: 96", "Number of lights: 2", "Successfully created",
                  "virtual light sampling : 1400 virtual lights in", "rendering : 122880 pixels in", "read render data :", "Total time:"]:
        k = out.find(token, pos)
        assert k >= 0, "stdout misses %r after offset %d:\n%s" % (token, pos, out)
        pos = k
    assert "Light 0:" not in out                       # this host does not echo the lights (CLSuperBidirectionalPathTracer.c:121-140)
    ref = oracle_fma.render("bidir", 384, 320, (1, 2, 3, 4), oracle_fma.load_scene_dir(d, "bidir"), n_vlp=700, want_accum=False, want_rng=False)
    tmp = os.path.join(d, "oracle_expected.ppm")
    oracle_fma.save_pam(tmp, ref["image"])
    assert open(os.path.join(d, "result.ppm"), "rb").read() == open(tmp, "rb").read()
    # the reference's own bytes with the separately rounded policy
    g = np.load(G)
    env["PT_ARITH"] = "separate"
    p = subprocess.run([BIN], cwd=d, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    raw = open(os.path.join(d, "result.ppm"), "rb").read()
    k = raw.index(b"ENDHDR\n") + 7
    assert hashlib.sha256(raw[k:]).digest() == g["img_s0_sha256"].tobytes()
    assert re.search(r"virtual light sampling : 1024 virtual lights", p.stdout)
