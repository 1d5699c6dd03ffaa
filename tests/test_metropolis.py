"""CLSuperMetropolisPathTracer(_vlpgrid): kernels lightTracer (seed paths) and MetropolisLightTracer in FIX mode.

As written the Metropolis kernels have no defined behaviour: VerifyIntersection hands TraceRay an uninitialised `float t` as its
running hit bound (metropolispathtracer.ocl:225-236, vlpgrid :239-242) and the host gives lightTracer the wrong buffer
(DESIGN.md section 7).  FIX mode = ONE patched line (`float t = 1e9;`, applied by oracle/Makefile to the text piped into the
compiler: libref_vlpgrid_fix.so) + the seed paths in their own buffer.  tests/golden/golden_metropolis.npz holds what the
reference's own (patched) kernels and its own Mutate function (through a probe: the kernel keeps the mutated path private)
produce; make_golden.py is the generating script.  Bar: bit-exact — seed-path lengths and vertices, mutated paths, VPL buffers —
for the oracle here and for CUDA (pt_launch_metropolis_lighttracer through the C ABI) on the GPU box.

What the goldens show about the program itself: every helper takes the RNG state by value, and the first pair of every
work-item is (seeds.x ^ seeds.z, seeds.y ^ seeds.w) * 2^-32 — with the host's 27-bit seeds always < 0.03125 — so GetRandomDirection
rejects it, the "second" direction of Mutate equals the first, and Mutate changes no path at all (0 of 1792 probed); seeds
outside that range (`extra*`) reach the other branches, and one set makes Mutate extend 11 of 512 paths."""
import os
import re
import subprocess

import numpy as np
import pytest

import opencl_montecarlo_path_tracing_b200 as pt
from conftest import ROOT, SEED_SETS

GOLDEN = np.load(os.path.join(ROOT, "tests", "golden", "golden_metropolis.npz"))
CASES = [(si, n, r) for si in (0, 1) for (n, r) in ((512, 8), (512, 0), (300, 1), (128, 40))]


def _same_path(a, b):
    n = int(b[16])
    return a[16] == b[16] and np.array_equal(a[:4 * n].reshape(-1, 4)[:, :3], b[:4 * n].reshape(-1, 4)[:, :3])


def _check(paths, vpl, gp, gv, what):
    assert np.array_equal(paths[:, 16], gp[:, 16]), what + ": seed-path lengths"
    assert all(_same_path(paths[i], gp[i]) for i in range(paths.shape[0])), what + ": seed-path vertices"
    assert np.array_equal(np.ascontiguousarray(vpl, np.float32).view(np.uint32), gv), what + ": VPL buffer"


@pytest.fixture(scope="module")
def metro_scene(oracle_sep, scene_dirs):
    return oracle_sep.load_scene_dir(scene_dirs["bidir"], "bidir")


@pytest.mark.parametrize("si,n,r", CASES)
def test_oracle_metropolis_kernels_equal_the_patched_reference(oracle_sep, metro_scene, si, n, r):
    paths, vpl = oracle_sep.metropolis_light_tracer(SEED_SETS[si], metro_scene, n, r)
    key = "s%d_n%d_r%d" % (si, n, r)
    _check(paths, vpl, GOLDEN[key + "_paths"], GOLDEN[key + "_vpl"], key)


def test_oracle_metropolis_lights_below_the_squares_and_wide_seeds(oracle_sep, metro_scene):
    sc = dict(metro_scene, lights=GOLDEN["below_lights"])
    for r in (0, 3, 8):
        paths, vpl = oracle_sep.metropolis_light_tracer(SEED_SETS[0], sc, 256, r)
        _check(paths, vpl, GOLDEN["below_n256_r%d_paths" % r], GOLDEN["below_n256_r%d_vpl" % r], "below r%d" % r)
    for ei, seeds in enumerate(GOLDEN["extra_seeds"]):
        paths, vpl = oracle_sep.metropolis_light_tracer(tuple(int(x) for x in seeds), metro_scene, 256, 8)
        _check(paths, vpl, GOLDEN["extra%d_paths" % ei], GOLDEN["extra%d_vpl" % ei], "extra%d" % ei)


def test_oracle_mutate_equals_the_references_own_function(oracle_sep, metro_scene):
    paths0 = GOLDEN["s0_n512_r0_paths"]
    for r in (1, 2, 8):
        g = GOLDEN["mutate_r%d" % r]
        for k, (gi, l) in enumerate(GOLDEN["mutate_which"]):
            p = oracle_sep.metropolis_mutate(SEED_SETS[0], metro_scene, int(gi), metro_scene["lights"][l][:3], paths0[gi + l * 512], r)
            assert _same_path(p, g[k]), (r, gi, l)
    changed = 0
    nl = metro_scene["lights"].shape[0]
    for ei, seeds in enumerate(GOLDEN["extra_seeds"]):
        seeds = tuple(int(x) for x in seeds)
        gp, gm = GOLDEN["extra%d_paths" % ei], GOLDEN["extra%d_mutated" % ei]
        k = 0
        for gi in range(0, 256):
            for l in range(nl):
                if ei == 4 or gi % 8 == 0:
                    p = oracle_sep.metropolis_mutate(seeds, metro_scene, gi, metro_scene["lights"][l][:3], gp[gi + l * 256], 8)
                    assert _same_path(p, gm[k]), (ei, gi, l)
                    changed += not _same_path(gm[k], gp[gi + l * 256])
                k += 1
    assert changed >= 10          # the seed set whose first pair is accepted and points down: Mutate really extends paths


@pytest.mark.gpu
@pytest.mark.parametrize("arith", ["separate", "fma"])
def test_cuda_metropolis_light_tracer_equals_reference_and_oracle(renderer, oracle_sep, oracle_fma, scene_dirs, arith):
    o = oracle_fma if arith == "fma" else oracle_sep
    osc = o.load_scene_dir(scene_dirs["bidir"], "bidir")
    scene = pt.load_scene_dir(scene_dirs["bidir"], "bidir")
    renderer.set_scene(scene)
    for (si, n, r) in CASES:
        paths, vpl = renderer.metropolis_light_tracer(SEED_SETS[si], n, r, arith=arith)
        op, ov = o.metropolis_light_tracer(SEED_SETS[si], osc, n, r)
        what = "s%d n%d r%d %s" % (si, n, r, arith)
        _check(paths, vpl, op, ov.view(np.uint32), what + " vs oracle")
        if arith == "separate":
            key = "s%d_n%d_r%d" % (si, n, r)
            _check(paths, vpl, GOLDEN[key + "_paths"], GOLDEN[key + "_vpl"], what + " vs the reference's bytes")
    for ei, seeds in enumerate(GOLDEN["extra_seeds"]):
        seeds = tuple(int(x) for x in seeds)
        paths, vpl = renderer.metropolis_light_tracer(seeds, 256, 8, arith=arith)
        op, ov = o.metropolis_light_tracer(seeds, osc, 256, 8)
        _check(paths, vpl, op, ov.view(np.uint32), "extra%d %s vs oracle" % (ei, arith))
        mp = renderer.read_metropolis_paths(mutated=True)
        om = np.array([o.metropolis_mutate(seeds, osc, gi, osc["lights"][l][:3], op[gi + l * 256], 8)
                       for l in range(osc["lights"].shape[0]) for gi in range(256)])
        assert all(_same_path(mp[i], om[i]) for i in range(mp.shape[0])), "extra%d %s: mutated paths" % (ei, arith)


@pytest.mark.gpu
def test_cuda_metropolis_vpls_feed_the_path_tracers(renderer, oracle_sep, scene_dirs):
    """the Metropolis VPL buffer through the VLP-grid path tracer (the program's own consumer), CUDA vs oracle"""
    osc = oracle_sep.load_scene_dir(scene_dirs["bidir"], "bidir")
    scene = pt.load_scene_dir(scene_dirs["bidir"], "bidir")
    renderer.set_scene(scene)
    _, vpl = renderer.metropolis_light_tracer(SEED_SETS[0], 512, 8, arith="separate")
    lo, hi = renderer.vlp_bounds()
    renderer.build_vlp_grid(pt.vlp_grid_dims(lo, hi, vpl.shape[0], 3.0))
    rows = (340, 372)
    res = renderer.render("vlpgrid", 512, 512, SEED_SETS[0], rows=rows, arith="separate", want_accum=True)
    ref = oracle_sep.render("vlpgrid", 512, 512, SEED_SETS[0], osc, vpls=vpl, rows=rows)
    assert np.array_equal(res.accum[rows[0]:rows[1]].view(np.uint32), ref["accum"][rows[0]:rows[1]].view(np.uint32))
    assert np.array_equal(res.image[rows[0]:rows[1]], ref["image"][rows[0]:rows[1]])


@pytest.mark.gpu
def test_cli_metropolis_vlpgrid_dropin(scene_dirs, oracle_fma):
    """bin/CLSuperMetropolisPathTracer_vlpgrid/CLSuperMetropolisPathTracer: the reference host's argv and stdout lines in its
    order (CLSuperMetropolisPathTracer.c:429-720), result.ppm == the oracle's FIX-mode pipeline for the printed seeds."""
    exe = os.path.join(ROOT, "opencl_montecarlo_path_tracing_b200", "bin", "CLSuperMetropolisPathTracer_vlpgrid", "CLSuperMetropolisPathTracer")
    d = scene_dirs["bidir"]
    env = dict(os.environ, PT_SEEDS="1,2,3,4")
    W, H, n_paths, rounds, mod = 192, 128, 300, 5, 2.5
    p = subprocess.run([exe, str(W), str(H), str(n_paths), str(rounds), str(mod)], cwd=d, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    out, pos = p.stdout, 0
    for token in ["Usage:", "[N_seedpaths_per_light] [mutation_rounds] [CELL_SIZE_MODIFIER]", "number of platforms:", "selected device", "Seeds: 1, 2, 3, 4",
                  "Processing image 192x128 with data size 98304 bytes", "Cam values:", "Number of triangles:", "Number of lights: 2",
                  "Mutation rounds: 5", "gws: 2560, lws: 256", "gws: 256, lws: 256", "VLPs bounding box values:", "VLPs grid size:",
                  "Successfully created render image result.ppm", "light paths random sampling : 600 random light paths",
                  "light paths metropolis sampling : 2400 virtual lights", "VLPs min/max reduction", "Read VLPs bounding box", "init VLPs grid :",
                  "rendering : 24576 pixels", "read render data : 98304 uchar", "Total time:"]:
        k = out.find(token, pos)
        assert k >= 0, "stdout misses %r after offset %d:\n%s" % (token, pos, out)
        pos = k
    osc = oracle_fma.load_scene_dir(d, "bidir")
    _, vpl = oracle_fma.metropolis_light_tracer((1, 2, 3, 4), osc, n_paths, rounds)
    ref = oracle_fma.render("vlpgrid", W, H, (1, 2, 3, 4), osc, vpls=vpl, modifier=mod, want_accum=False, want_rng=False)
    g = ref["vlp_grid"]
    m = re.search(r"VLPs grid size: (\d+) x (\d+) x (\d+)", out)
    assert [int(x) for x in m.groups()] == [int(x) for x in g["res"][:3]]
    lo, hi = oracle_fma.vlp_bounds(vpl)
    assert ("vmax: %f %f %f, vmin: %f %f %f" % (hi[0], hi[1], hi[2], lo[0], lo[1], lo[2])) in out
    tmp = os.path.join(d, "oracle_expected_metro.ppm")
    oracle_fma.save_pam(tmp, ref["image"])
    assert open(os.path.join(d, "result.ppm"), "rb").read() == open(tmp, "rb").read()
