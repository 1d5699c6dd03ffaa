"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Bar: BIT-EXACT — image bytes, the float accumulation buffer, the final per-work-item
RNG state and the work counters — for both arithmetic policies:
    arith=fma       vs oracle built -DPT_CONTRACT=1
    arith=separate  vs oracle built -DPT_CONTRACT=0 (itself byte-identical to the reference .ocl
                    compiled for the CPU, tests/test_oracle_vs_ref.py)
"""
import numpy as np
import pytest

import opencl_montecarlo_path_tracing_b200 as pt
from conftest import SEED_SETS

pytestmark = pytest.mark.gpu

# row windows of the 512x512 default view that contain every kind of content
WINDOWS = {"sky+mesh+sphere+square": (112, 144), "far squares+horizon": (196, 260), "floor+sphere+shadow": (340, 372),
           "bottom": (496, 512)}


def _oracle_scene(o, d, variant):
    return o.load_scene_dir(d, variant)


def _compare(res, ref, what):
    assert np.array_equal(res.rng_state, ref["rng_state"]), "%s: RNG state differs in %d work-items" % (
        what, int((res.rng_state != ref["rng_state"]).any(axis=1).sum()))
    a, b = res.accum.view(np.uint32), ref["accum"].view(np.uint32)
    bad = int((a != b).any(axis=2).sum())
    assert bad == 0, "%s: accumulation buffer differs in %d pixels (max abs %g)" % (
        what, bad, float(np.abs(res.accum - ref["accum"]).max()))
    assert np.array_equal(res.image, ref["image"]), "%s: image bytes differ" % what
    for k in ("samples", "rays", "shadow_rays", "tri_tests", "cells_visited", "prim_tests"):
        assert res.counters[k] == ref["counters"][k], "%s: counter %s %d != %d" % (what, k, res.counters[k], ref["counters"][k])


@pytest.mark.parametrize("variant", ["base", "lmem", "nodof", "grid"])
@pytest.mark.parametrize("arith", ["fma", "separate"])
@pytest.mark.parametrize("kernel", ["mega", "persistent"])
def test_bit_exact_windows(renderer, scene_dirs, oracle_fma, oracle_sep, variant, arith, kernel):
    o = oracle_fma if arith == "fma" else oracle_sep
    d = scene_dirs[variant]
    scene = pt.load_scene_dir(d, variant)
    renderer.set_scene(scene)
    osc = _oracle_scene(o, d, variant)
    if variant == "grid":
        renderer.build_grid(pt.grid_dims(scene))
    W = H = 512
    for name, rows in WINDOWS.items():
        res = renderer.render(variant, W, H, SEED_SETS[0], rows=rows, arith=arith, kernel=kernel, want_accum=True, want_rng=True)
        ref = o.render(variant, W, H, SEED_SETS[0], osc, rows=rows)
        r0, r1 = rows
        if variant == "nodof":
            # rng states are per sample: compare only the rendered rows' work-items
            st = res.rng_state.reshape(8 * H, 8 * W, 4)[8 * r0:8 * r1]
            rst = ref["rng_state"].reshape(8 * H, 8 * W, 4)[8 * r0:8 * r1]
        else:
            st = res.rng_state.reshape(H, W, 4)[r0:r1]
            rst = ref["rng_state"].reshape(H, W, 4)[r0:r1]
        res.rng_state, ref["rng_state"] = st.reshape(-1, 4), rst.reshape(-1, 4)
        res.accum, ref["accum"] = res.accum[r0:r1], ref["accum"][r0:r1]
        res.image, ref["image"] = res.image[r0:r1], ref["image"][r0:r1]
        _compare(res, ref, "%s/%s/%s/%s" % (variant, arith, kernel, name))


@pytest.mark.parametrize("variant", ["base", "lmem", "nodof", "grid"])
def test_scene_mem_and_seeds(renderer, scene_dirs, oracle_fma, variant):
    """constant-memory and shared-memory scene staging give the same bits; second seed set; non-square size."""
    d = scene_dirs[variant]
    scene = pt.load_scene_dir(d, variant)
    renderer.set_scene(scene)
    if variant == "grid":
        renderer.build_grid(pt.grid_dims(scene))
    W, H, rows = 640, 360, (300, 332)
    ref = oracle_fma.render(variant, W, H, SEED_SETS[1], _oracle_scene(oracle_fma, d, variant), rows=rows)
    for kernel in ("mega", "persistent"):
        for mem in ("const", "smem"):
            res = renderer.render(variant, W, H, SEED_SETS[1], rows=rows, scene_mem=mem, kernel=kernel, want_accum=True)
            assert np.array_equal(res.image[rows[0]:rows[1]], ref["image"][rows[0]:rows[1]]), (kernel, mem)
            assert np.array_equal(res.accum[rows[0]:rows[1]].view(np.uint32), ref["accum"][rows[0]:rows[1]].view(np.uint32)), (kernel, mem)


@pytest.mark.parametrize("variant", ["base", "lmem", "nodof", "grid"])
def test_full_frame_bit_exact(renderer, scene_dirs, oracle_fma, variant):
    """Whole 512x512x64 default frame of every variant (default kernel choice): image, accumulation buffer,
    RNG state of every work-item and all work counters equal the oracle's."""
    d = scene_dirs[variant]
    scene = pt.load_scene_dir(d, variant)
    renderer.set_scene(scene)
    if variant == "grid":
        renderer.build_grid(pt.grid_dims(scene))
    res = renderer.render(variant, 512, 512, SEED_SETS[1], want_accum=True, want_rng=True)
    ref = oracle_fma.render(variant, 512, 512, SEED_SETS[1], _oracle_scene(oracle_fma, d, variant))
    _compare(res, ref, "%s/full frame" % variant)


def test_torus_mesh_and_spp_extension(renderer, scene_dirs, oracle_fma):
    """config 3 ingredients: torus.txt as triangles.txt, spp != 64 (continued stream, scale 224/spp)."""
    d = scene_dirs["torus"]
    scene = pt.load_scene_dir(d, "base")
    assert scene.ntriangles == 32
    renderer.set_scene(scene)
    W, H, rows = 512, 512, (150, 182)
    osc = _oracle_scene(oracle_fma, d, "base")
    for spp in (16, 256):
        res = renderer.render("base", W, H, SEED_SETS[0], rows=rows, spp=spp, want_accum=True, want_rng=True)
        ref = oracle_fma.render("base", W, H, SEED_SETS[0], osc, rows=rows, spp=spp)
        assert np.array_equal(res.image[rows[0]:rows[1]], ref["image"][rows[0]:rows[1]])
        assert np.array_equal(res.rng_state.reshape(H, W, 4)[rows[0]:rows[1]], ref["rng_state"].reshape(H, W, 4)[rows[0]:rows[1]])


def test_rng_probe_matches_known_answers(renderer, oracle_sep):
    for seeds, gid in (((1, 2, 3, 4), 0), ((1, 2, 3, 4), 12345), ((123456789, 42, 7, 99999), 262143)):
        f, st = renderer.probe_rng(seeds, gid, 64)
        rf, ru, rst = oracle_sep.rng_kat(seeds, gid, 64)
        assert np.array_equal(f.view(np.uint32), rf.view(np.uint32))
        assert np.array_equal(st, rst)


def test_full_frame_matches_oracle_and_properties(renderer, scene_dirs, oracle_fma):
    """Whole 256x256 frame of the NoDoF variant (the benchmark workload) + size-independent properties."""
    d = scene_dirs["nodof"]
    scene = pt.load_scene_dir(d, "nodof")
    renderer.set_scene(scene)
    res = renderer.render("nodof", 256, 256, SEED_SETS[0], want_accum=True)
    ref = oracle_fma.render("nodof", 256, 256, SEED_SETS[0], _oracle_scene(oracle_fma, d, "nodof"), want_rng=False)
    assert np.array_equal(res.image, ref["image"])
    assert res.counters["rays"] == ref["counters"]["rays"]
    # properties: alpha is 255 everywhere, determinism, tiles compose to the full frame bit-for-bit
    assert (res.image[..., 3] == 255).all()
    again = renderer.render("nodof", 256, 256, SEED_SETS[0])
    assert np.array_equal(again.image, res.image)
    top = renderer.render("nodof", 256, 256, SEED_SETS[0], rows=(0, 100)).image
    bot = renderer.render("nodof", 256, 256, SEED_SETS[0], rows=(100, 256)).image
    assert np.array_equal(np.concatenate([top[:100], bot[100:]]), res.image)
    # interleaved stripes of 2 "ranks" also compose exactly (the multi-GPU sharding)
    parts = [renderer.render("nodof", 256, 256, SEED_SETS[0], interleave=8, rank=r, nranks=2).image.astype(np.int32) for r in range(2)]
    assert np.array_equal((parts[0] + parts[1]).astype(np.uint8), res.image)
