"""Parity at the sizes bench.py times (BASELINE.json configs 3, 4 and 5), CUDA through the C ABI vs the CPU oracle.

Bar: BIT-EXACT (image bytes, float accumulation buffer, final RNG state of every work-item, work counters), on row
windows the oracle finishes in seconds:
  * config 4/5 scene — gen_mesh.soup(1 << 20): 1,048,576 triangles, 128^3 grid, 32-bit cell ids, the `BIG` grid
    instantiation of the megakernel — at 1920x1080x256 and 3840x2160 (64 spp windows + one row at the full 4096 spp),
    device-built CSR == oracle CSR at 1 M triangles;
  * config 3 — triangles.txt and torus.txt at 1920x1080x1024 with the per-cluster cull compiled in (the `BIG` brute-force
    instantiation bench.py's base_1920x1080x1024 / torus_1920x1080x1024 run), cull on == cull off.
Follows CLSuperPathTracer_trianglegrid/pathtracer.ocl:102-201 (grid TraceRay) and CLSuperPathTracer/pathtracer.ocl:48-137.
"""
import numpy as np
import pytest

import opencl_montecarlo_path_tracing_b200 as pt
from conftest import SEED_SETS

pytestmark = pytest.mark.gpu

SOUP_N = 1 << 20


@pytest.fixture(scope="module")
def soup1m():
    import gen_mesh
    tris = gen_mesh.soup(SOUP_N)
    lo, hi = gen_mesh.bbox_like_reference(tris)
    sph = np.array([1024, 0, 0, 0, 145, 0, 0, 2048, 0], np.int32)
    sq = np.array([4096, 0, 0, 0, 0, 0, 129, 0, 8192], np.int32)
    lights = np.array([[10, 4, 10, 400], [15, 2, 7, 300]], np.float32)
    scene = pt.Scene(sph, sq, tris, lights, lo, hi)
    osc = {"spheres": sph, "squares": sq, "triangles": tris, "lights": lights, "box_min": lo, "box_max": hi}
    return scene, osc


@pytest.fixture(scope="module")
def soup_renderer(soup1m):
    scene, _ = soup1m
    r = pt.Renderer(device=0)
    r.set_scene(scene)
    g = pt.grid_dims(scene)
    assert list(g.res[:3]) == [128, 128, 128]
    r.build_grid(g)
    yield r, g
    r.close()


_OGRID = {}


def _ogrid(o, osc, g):
    if o.contract not in _OGRID:
        res, cell = np.array(g.res[:], np.int32), np.array(g.cell_size[:], np.float32)
        ores, ocell = o.grid_dims(osc["box_min"], osc["box_max"], SOUP_N, 3.0)
        assert np.array_equal(res[:3], ores[:3]) and np.array_equal(cell[:3].view(np.uint32), ocell[:3].view(np.uint32))
        _OGRID[o.contract] = {"box_min": osc["box_min"], "box_max": osc["box_max"], "res": res, "cell_size": cell,
                              "csr": o.build_grid(osc["triangles"], osc["box_min"], res, cell)}
    return _OGRID[o.contract]


def _check_window(r, o, osc, grid, W, H, spp, rows, what, **kw):
    res = r.render("grid", W, H, SEED_SETS[0], rows=rows, spp=spp, want_accum=True, want_rng=True, **kw)
    ref = o.render("grid", W, H, SEED_SETS[0], osc, spp=spp, rows=rows, grid=grid)
    r0, r1 = rows
    a, b = res.rng_state.reshape(H, W, 4)[r0:r1], ref["rng_state"].reshape(H, W, 4)[r0:r1]
    assert np.array_equal(a, b), "%s: RNG state differs in %d pixels" % (what, int((a != b).any(axis=2).sum()))
    a, b = res.accum[r0:r1].view(np.uint32), ref["accum"][r0:r1].view(np.uint32)
    assert np.array_equal(a, b), "%s: accumulation differs in %d pixels" % (what, int((a != b).any(axis=2).sum()))
    assert np.array_equal(res.image[r0:r1], ref["image"][r0:r1]), what
    for k in ("samples", "rays", "shadow_rays", "tri_tests", "cells_visited", "prim_tests"):
        assert res.counters[k] == ref["counters"][k], "%s: counter %s %d != %d" % (what, k, res.counters[k], ref["counters"][k])


def test_soup1m_device_grid_equals_oracle_grid(soup_renderer, soup1m, oracle_fma):
    r, g = soup_renderer
    start, refs = r.read_grid_csr()
    ostart, orefs = _ogrid(oracle_fma, soup1m[1], g)["csr"]
    assert start.shape == ostart.shape and np.array_equal(start, ostart)
    assert np.array_equal(refs, orefs)
    n = np.diff(start.astype(np.int64))
    assert n.max() <= 62 and refs.size > 3_000_000


@pytest.mark.parametrize("arith", ["fma", "separate"])
@pytest.mark.parametrize("kernel", ["auto", "mega", "persistent", "grid_pool", "grid_queue", "grid_async"])
def test_config4_soup1m_1920x1080x256_windows(soup_renderer, soup1m, oracle_fma, oracle_sep, arith, kernel):
    r, g = soup_renderer
    o = oracle_fma if arith == "fma" else oracle_sep
    grid = _ogrid(o, soup1m[1], g)
    # windows of 4 rows spread over the frame: sky + box top, box interior, box + floor, floor at the bottom
    windows = [(100, 104), (420, 424), (700, 704), (1072, 1076)] if kernel not in ("persistent",) else [(420, 424), (1072, 1076)]
    for rows in windows:
        _check_window(r, o, soup1m[1], grid, 1920, 1080, 256, rows, "config 4 %s/%s rows %s" % (kernel, arith, rows), arith=arith, kernel=kernel)


@pytest.mark.parametrize("arith", ["fma", "separate"])
def test_config5_soup1m_3840x2160_windows(soup_renderer, soup1m, oracle_fma, oracle_sep, arith):
    r, g = soup_renderer
    o = oracle_fma if arith == "fma" else oracle_sep
    grid = _ogrid(o, soup1m[1], g)
    for rows in [(640, 644), (1500, 1504)]:
        _check_window(r, o, soup1m[1], grid, 3840, 2160, 64, rows, "config 5 @64 spp %s rows %s" % (arith, rows), arith=arith)


def test_config5_one_row_at_the_full_4096_spp(soup_renderer, soup1m, oracle_fma):
    r, g = soup_renderer
    grid = _ogrid(oracle_fma, soup1m[1], g)
    _check_window(r, oracle_fma, soup1m[1], grid, 3840, 2160, 4096, (1111, 1112), "config 5 row 1111 @4096 spp")


def test_config5_stripes_of_8_ranks_compose_the_frame_window(soup_renderer, soup1m):
    """The multi-GPU sharding of config 5 (8-row stripes dealt to 8 ranks) on one device: the sum of the 8 ranks' float
    buffers over a 64-row band equals the unsharded render bit for bit (each pixel is written by exactly one rank)."""
    r, _ = soup_renderer
    W, H, spp, rows = 3840, 2160, 8, (1024, 1088)
    full = r.render("grid", W, H, SEED_SETS[0], rows=rows, spp=spp, want_accum=True).accum[rows[0]:rows[1]]
    acc = np.zeros_like(full)
    owners = np.zeros(full.shape[:2], np.int32)
    for rank in range(8):
        part = r.render("grid", W, H, SEED_SETS[0], rows=rows, spp=spp, want_accum=True, interleave=8, rank=rank, nranks=8).accum[rows[0]:rows[1]]
        owners += (part[..., 3] != 0)
        acc += part
    assert (owners == 1).all()
    assert np.array_equal(acc.view(np.uint32), full.view(np.uint32))


@pytest.mark.parametrize("mesh", ["base", "torus"])
@pytest.mark.parametrize("arith", ["fma", "separate"])
def test_config3_1920x1080x1024_windows_with_cluster_cull(renderer, scene_dirs, oracle_fma, oracle_sep, mesh, arith):
    """config 3 at its true size: windows through the mesh, the spheres/squares and the floor, rendered by the `BIG`
    megakernel instantiation (per-cluster cull compiled in; cluster_cull="on" forces what full frames > 400 k pixels get)."""
    o = oracle_fma if arith == "fma" else oracle_sep
    d = scene_dirs[mesh]
    scene = pt.load_scene_dir(d, "base")
    renderer.set_scene(scene)
    osc = o.load_scene_dir(d, "base")
    W, H, spp = 1920, 1080, 1024
    for rows in [(150, 152), (330, 331), (700, 701)]:       # mesh (rows 124-200), squares+spheres, floor
        ref = o.render("base", W, H, SEED_SETS[0], osc, spp=spp, rows=rows)
        r0, r1 = rows
        for kernel, cc in (("mega", "on"), ("mega", "off"), ("persistent", "auto")):
            res = renderer.render("base", W, H, SEED_SETS[0], rows=rows, spp=spp, arith=arith, kernel=kernel, cluster_cull=cc,
                                  want_accum=True, want_rng=True)
            what = "config 3 %s %s %s/%s rows %s" % (mesh, arith, kernel, cc, rows)
            assert np.array_equal(res.rng_state.reshape(H, W, 4)[r0:r1], ref["rng_state"].reshape(H, W, 4)[r0:r1]), what
            assert np.array_equal(res.accum[r0:r1].view(np.uint32), ref["accum"][r0:r1].view(np.uint32)), what
            assert np.array_equal(res.image[r0:r1], ref["image"][r0:r1]), what
            for k in ("samples", "rays", "shadow_rays", "tri_tests", "prim_tests"):
                assert res.counters[k] == ref["counters"][k], (what, k)
            if cc == "on":
                assert res.counters["tri_tests_executed"] < res.counters["tri_tests"]
