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
@pytest.mark.parametrize("kernel", ["mega", "persistent", "wavefront"])
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


def _soup_scene(n, seed, box):
    import gen_mesh
    tris = gen_mesh.soup(n, seed=seed, box_size=box)
    lo, hi = gen_mesh.bbox_like_reference(tris)
    base = pt.Scene(np.array([1024, 0, 0, 0, 145, 0, 0, 2048, 0], np.int32), np.array([4096, 0, 0, 0, 0, 0, 129, 0, 8192], np.int32),
                    tris, np.array([[10, 4, 10, 400], [15, 2, 7, 300]], np.float32), lo, hi)
    osc = {"spheres": base.spheres, "squares": base.squares, "triangles": tris, "lights": base.lights, "box_min": lo, "box_max": hi}
    return base, osc


@pytest.mark.parametrize("n,box,kernel", [(30000, 24.0, "mega"), (30000, 24.0, "persistent"), (30000, 24.0, "wavefront"),
                                          (30000, 24.0, "grid_tma"), (30000, 24.0, "grid_stream"), (70000, 30.0, "mega"),
                                          (70000, 30.0, "grid_stream"), (30000, 24.0, "grid_pool"), (70000, 30.0, "grid_pool"),
                                          (30000, 24.0, "grid_queue"), (70000, 30.0, "grid_queue"),
                                          (30000, 24.0, "grid_async"), (70000, 30.0, "grid_async")])
def test_grid_synthetic_soup_bit_exact(renderer, oracle_fma, n, box, kernel):
    """Uniform-grid traversal on a seeded triangle soup (the config-4/5 generator at test size); 70000
    triangles exceed the reference's 16-bit cell ids (wide ids).  The camera sits inside the box."""
    scene, osc = _soup_scene(n, 11, box)
    renderer.set_scene(scene)
    g = pt.grid_dims(scene)
    renderer.build_grid(g)
    # device-built grid == oracle's deterministic binning
    start, refs = renderer.read_grid_csr()
    ostart, orefs = oracle_fma.build_grid(osc["triangles"], osc["box_min"], np.array(g.res[:]), np.array(g.cell_size[:], np.float32))
    assert np.array_equal(start, ostart) and np.array_equal(refs, orefs)
    W, H, rows = 512, 512, (250, 266)
    res = renderer.render("grid", W, H, SEED_SETS[0], rows=rows, kernel=kernel, want_accum=True, want_rng=True)
    ref = oracle_fma.render("grid", W, H, SEED_SETS[0], osc, rows=rows)
    r0, r1 = rows
    res.rng_state, ref["rng_state"] = res.rng_state.reshape(H, W, 4)[r0:r1].reshape(-1, 4), ref["rng_state"].reshape(H, W, 4)[r0:r1].reshape(-1, 4)
    res.accum, ref["accum"] = res.accum[r0:r1], ref["accum"][r0:r1]
    res.image, ref["image"] = res.image[r0:r1], ref["image"][r0:r1]
    _compare(res, ref, "grid soup %d %s" % (n, kernel))
    assert res.counters["cells_visited"] > 0 and res.counters["tri_tests"] > 0


@pytest.mark.parametrize("kernel", ["grid_tma", "grid_stream", "grid_pool", "grid_queue", "grid_async"])
def test_grid_tma_default_scene_and_dense_cells(renderer, scene_dirs, oracle_fma, kernel):
    """TMA-staged warp-per-ray traversal and cell-granular regeneration: default grid scene, and CELL_SIZE_MODIFIER 0.02
    (one fat cell capped at 62)."""
    d = scene_dirs["grid"]
    scene = pt.load_scene_dir(d, "grid")
    renderer.set_scene(scene)
    osc = _oracle_scene(oracle_fma, d, "grid")
    W = H = 512
    for mod, rows in ((3.0, (112, 144)), (0.02, (120, 136))):
        g = pt.grid_dims(scene, mod)
        renderer.build_grid(g)
        res = renderer.render("grid", W, H, SEED_SETS[0], rows=rows, kernel=kernel, want_accum=True, want_rng=True)
        ref = oracle_fma.render("grid", W, H, SEED_SETS[0], osc, rows=rows, modifier=mod, grid=None)
        r0, r1 = rows
        assert np.array_equal(res.image[r0:r1], ref["image"][r0:r1]), mod
        assert np.array_equal(res.accum[r0:r1].view(np.uint32), ref["accum"][r0:r1].view(np.uint32)), mod
        assert np.array_equal(res.rng_state.reshape(H, W, 4)[r0:r1], ref["rng_state"].reshape(H, W, 4)[r0:r1]), mod
        for k in ("rays", "shadow_rays", "cells_visited", "tri_tests"):
            assert res.counters[k] == ref["counters"][k], (mod, k)


def test_grid_build_matches_reference_cells(renderer, scene_dirs):
    """pt_read_grid_cells exports the reference's 128-byte Cell layout; contents equal the golden cells
    produced by the reference's own initTrianglesGrid kernel (as sets; ours are in triangle-id order)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_grid.npz"))
    for name, key, mod in (("default", "grid", 3.0), ("torus", "torus", 3.0), ("torus_fine", "torus", 40.0)):
        scene = pt.load_scene_dir(scene_dirs[key], "grid")
        renderer.set_scene(scene)
        renderer.build_grid(pt.grid_dims(scene, mod))
        cells = renderer.read_grid_cells()
        nels = cells[:, :4].copy().view(np.uint32).reshape(-1)
        ids = cells[:, 4:].copy().view(np.uint16).reshape(-1, 62)
        gn = g[name + "_nels"]
        assert np.array_equal(nels, np.minimum(gn, 62))
        for c in np.nonzero(gn <= 62)[0]:
            assert np.array_equal(ids[c, :gn[c]], g[name + "_ids"][c][:gn[c]])


@pytest.mark.parametrize("kernel", ["mega", "persistent", "wavefront"])
def test_many_analytic_primitives(renderer, scene_dirs, oracle_fma, kernel):
    """More than 8 spheres / squares leaves the parameter-resident fast path (lists in the scene block)."""
    scene = pt.load_scene_dir(scene_dirs["lmem"], "lmem")
    scene.spheres = np.array([1024 + 64 + 4, 2, 0, 16384, 145, 0, 8 + 512, 2048, 65536], np.int32)      # 12 spheres
    scene.squares = np.array([4096, 1, 32, 0, 2 + 1024, 0, 129, 16, 8192 + 8], np.int32)               # 10 squares
    renderer.set_scene(scene)
    osc = {"spheres": scene.spheres, "squares": scene.squares, "triangles": scene.triangles, "lights": scene.lights}
    W = H = 512
    for rows in ((120, 136), (340, 356)):
        res = renderer.render("lmem", W, H, SEED_SETS[0], rows=rows, kernel=kernel, want_accum=True, want_rng=True)
        ref = oracle_fma.render("lmem", W, H, SEED_SETS[0], osc, rows=rows)
        r0, r1 = rows
        assert np.array_equal(res.accum[r0:r1].view(np.uint32), ref["accum"][r0:r1].view(np.uint32))
        assert np.array_equal(res.rng_state.reshape(H, W, 4)[r0:r1], ref["rng_state"].reshape(H, W, 4)[r0:r1])
        assert res.counters["prim_tests"] == ref["counters"]["prim_tests"]


def test_full_size_flavours_agree(renderer, scene_dirs):
    """At BASELINE.json's full image size (1920x1080) the oracle is too slow to render whole frames, so the
    size-independent property is checked instead: every kernel flavour, both scene memories and a 3-way row
    tiling produce bit-identical accumulation buffers (spp reduced to 16; work is identical per sample)."""
    W, H, spp = 1920, 1080, 16
    for variant, flavours in (("base", ("mega", "persistent", "wavefront")), ("nodof", ("mega", "persistent", "wavefront")),
                              ("grid", ("mega", "persistent", "wavefront", "grid_tma", "grid_stream", "grid_pool"))):
        scene = pt.load_scene_dir(scene_dirs[variant], variant)
        renderer.set_scene(scene)
        if variant == "grid":
            renderer.build_grid(pt.grid_dims(scene))
        s = 64 if variant == "nodof" else spp
        ref = renderer.render(variant, W, H, SEED_SETS[0], spp=s, kernel=flavours[0], want_accum=True)
        assert (ref.image[..., 3] == 255).all()
        for k in flavours[1:]:
            got = renderer.render(variant, W, H, SEED_SETS[0], spp=s, kernel=k, want_accum=True)
            assert np.array_equal(got.accum.view(np.uint32), ref.accum.view(np.uint32)), (variant, k)
            assert got.counters["rays"] == ref.counters["rays"]
        other_mem = renderer.render(variant, W, H, SEED_SETS[0], spp=s, kernel=flavours[0], scene_mem="const", want_accum=True)
        assert np.array_equal(other_mem.accum.view(np.uint32), ref.accum.view(np.uint32))
        parts = [renderer.render(variant, W, H, SEED_SETS[0], spp=s, rows=r, want_accum=True).accum for r in ((0, 300), (300, 777), (777, 1080))]
        tiled = np.concatenate([parts[0][:300], parts[1][300:777], parts[2][777:]])
        assert np.array_equal(tiled.view(np.uint32), ref.accum.view(np.uint32)), variant


@pytest.mark.parametrize("arith", ["fma", "separate"])
def test_conservative_culls_change_nothing(renderer, scene_dirs, arith):
    """Mesh bounding-sphere cull, analytic bounding-box cull (incl. their distance-proportional margins for the
    reference's far-field rounding): with and without them the float accumulation buffer is bit-identical,
    on whole frames that include the horizon, at two image sizes."""
    for variant, sdir in (("base", "base"), ("lmem", "lmem"), ("nodof", "nodof"), ("grid", "grid"), ("base", "torus"), ("bidir", "bidir")):
        scene = pt.load_scene_dir(scene_dirs[sdir], variant)
        renderer.set_scene(scene)
        if variant == "grid":
            renderer.build_grid(pt.grid_dims(scene))
        if variant == "bidir":
            renderer.light_tracer(SEED_SETS[1], 512, arith=arith)
        # frames above 400 k pixels additionally run the per-cluster triangle cull (8 consecutive records per cluster)
        for (W, H, spp) in ((512, 512, 64), (1920, 1080, 64 if variant == "nodof" else 8)):
            a = renderer.render(variant, W, H, SEED_SETS[1], spp=spp, arith=arith, want_accum=True, cull=True)
            b = renderer.render(variant, W, H, SEED_SETS[1], spp=spp, arith=arith, want_accum=True, cull=False)
            bad = int((a.accum.view(np.uint32) != b.accum.view(np.uint32)).any(axis=2).sum())
            assert bad == 0, "%s %dx%d %s: %d pixels change when the conservative culls are enabled" % (variant, W, H, arith, bad)
            assert a.counters["rays"] == b.counters["rays"]
            if variant in ("base", "lmem", "bidir"):
                assert a.counters["tri_tests_executed"] < b.counters["tri_tests_executed"] == b.counters["tri_tests"]
