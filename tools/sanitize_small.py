#!/usr/bin/env python3
"""Small frames through every variant / kernel flavour / extension, meant to be run under
`compute-sanitizer --tool memcheck` (and racecheck) on the GPU box: out-of-bounds or misaligned accesses abort it."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scenes"))
import numpy as np  # noqa: E402
import gen_mesh  # noqa: E402
import write_scenes  # noqa: E402
import opencl_montecarlo_path_tracing_b200 as pt  # noqa: E402

S = (1, 2, 3, 4)
W, H, ROWS = 512, 512, (348, 352)
n = 0
with pt.Renderer(0) as r, tempfile.TemporaryDirectory() as tmp:
    for v, kernels in (("base", ("mega", "persistent", "wavefront")), ("lmem", ("mega", "persistent", "wavefront")),
                       ("nodof", ("mega", "persistent", "wavefront")), ("grid", ("mega", "persistent", "wavefront", "grid_tma", "grid_stream", "grid_pool", "grid_queue", "grid_async")),
                       ("bidir", ("mega",)), ("vlpgrid", ("mega",))):
        d = os.path.join(tmp, v)
        sv = "bidir" if v == "vlpgrid" else v
        write_scenes.write_variant(sv, d)
        scene = pt.load_scene_dir(d, sv)
        r.set_scene(scene)
        if v == "grid":
            r.build_grid(pt.grid_dims(scene))
        if v == "bidir":
            r.light_tracer(S, 512); r.light_tracer(S, 33); r.read_vpls(); r.set_vpls(np.zeros((0, 4), np.float32)); r.light_tracer(S, 512)
        if v == "vlpgrid":
            r.metropolis_light_tracer(S, 64, 2); r.read_metropolis_paths(mutated=True)
            r.light_tracer(S, 512)
            lo, hi = r.vlp_bounds()
            r.build_vlp_grid(pt.vlp_grid_dims(lo, hi, 1024, 3.0))
        for k in kernels:
            if v in ("base", "lmem", "grid") and k == "mega":
                r.render(v, W, H, S, rows=ROWS, kernel=k, dead_rays="trace", want_accum=True); n += 1
                if v != "grid":
                    r.render(v, W, H, S, rows=(112, 116), kernel="spec", want_accum=True, want_rng=True); n += 1
            for mem in ("const", "smem"):
                r.render(v, W, H, S, rows=ROWS, kernel=k, scene_mem=mem, want_accum=True, want_rng=True); n += 1
            if v not in ("nodof", "vlpgrid"):
                r.render(v, W, H, S, rows=ROWS, kernel=k, sample_block=1, sample_blocks=4, want_accum=True); n += 1
                r.render(v, 70, 45, S, kernel=k, interleave=8, rank=1, nranks=3, want_accum=True); n += 1
    tris = gen_mesh.soup(20000, box_size=20.0)
    lo, hi = gen_mesh.bbox_like_reference(tris)
    scene = pt.Scene(scene.spheres, scene.squares, tris, scene.lights, lo, hi)
    r.set_scene(scene)
    r.build_grid(pt.grid_dims(scene))
    for k in ("mega", "persistent", "grid_tma", "grid_stream", "grid_pool", "grid_queue", "grid_async"):
        r.render("grid", W, H, S, rows=(250, 254), kernel=k); n += 1
    print("selftest", r.selftest_fastmath(1 << 16))
print("sanitize_small: %d launches completed" % n)
