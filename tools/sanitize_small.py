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
                       ("nodof", ("mega", "persistent", "wavefront")), ("grid", ("mega", "persistent", "wavefront", "grid_tma", "grid_stream")),
                       ("bidir", ("mega",))):
        d = os.path.join(tmp, v)
        write_scenes.write_variant(v, d)
        scene = pt.load_scene_dir(d, v)
        r.set_scene(scene)
        if v == "grid":
            r.build_grid(pt.grid_dims(scene))
        if v == "bidir":
            r.light_tracer(S, 512); r.light_tracer(S, 33); r.read_vpls(); r.set_vpls(np.zeros((0, 4), np.float32)); r.light_tracer(S, 512)
        for k in kernels:
            for mem in ("const", "smem"):
                r.render(v, W, H, S, rows=ROWS, kernel=k, scene_mem=mem, want_accum=True, want_rng=True); n += 1
            if v != "nodof":
                r.render(v, W, H, S, rows=ROWS, kernel=k, sample_block=1, sample_blocks=4, want_accum=True); n += 1
                r.render(v, 70, 45, S, kernel=k, interleave=8, rank=1, nranks=3, want_accum=True); n += 1
    tris = gen_mesh.soup(20000, box_size=20.0)
    lo, hi = gen_mesh.bbox_like_reference(tris)
    scene = pt.Scene(scene.spheres, scene.squares, tris, scene.lights, lo, hi)
    r.set_scene(scene)
    r.build_grid(pt.grid_dims(scene))
    for k in ("mega", "persistent", "grid_tma", "grid_stream"):
        r.render("grid", W, H, S, rows=(250, 254), kernel=k); n += 1
    print("selftest", r.selftest_fastmath(1 << 16))
print("sanitize_small: %d launches completed" % n)
