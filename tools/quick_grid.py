#!/usr/bin/env python3
"""Timing of the synthetic-soup grid configs (BASELINE configs 4/5 ingredients) on the GPU box.
env: QG_N (triangles, default 1048576) QG_W QG_H QG_SPP QG_BOX QG_KERNELS"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scenes"))
import numpy as np  # noqa: E402
import gen_mesh  # noqa: E402
import opencl_montecarlo_path_tracing_b200 as pt  # noqa: E402

n = int(os.environ.get("QG_N", str(1 << 20)))
W = int(os.environ.get("QG_W", "1920")); H = int(os.environ.get("QG_H", "1080")); spp = int(os.environ.get("QG_SPP", "16"))
box = float(os.environ.get("QG_BOX", "60"))
tris = gen_mesh.soup(n, box_size=box)
lo, hi = gen_mesh.bbox_like_reference(tris)
scene = pt.Scene(np.array([1024, 0, 0, 0, 145, 0, 0, 2048, 0], np.int32), np.array([4096, 0, 0, 0, 0, 0, 129, 0, 8192], np.int32),
                 tris, np.array([[10, 4, 10, 400], [15, 2, 7, 300]], np.float32), lo, hi)
with pt.Renderer(0) as r:
    t0 = time.time(); r.set_scene(scene); t1 = time.time()
    g = pt.grid_dims(scene, float(os.environ.get("QG_MOD", "3.0")))
    ms = r.build_grid(g)
    print("set_scene %.1f ms, grid %dx%dx%d built in %.2f ms (device)" % ((t1 - t0) * 1e3, g.res[0], g.res[1], g.res[2], ms), flush=True)
    for kernel in os.environ.get("QG_KERNELS", "mega,persistent,grid_tma").split(","):
        best = 1e9
        for it in range(3):
            res = r.render("grid", W, H, (1, 2, 3, 4), spp=spp, kernel=kernel, read_image=False, dead_rays=os.environ.get("QG_DEAD", "auto"))
            best = min(best, res.ms)
        c = res.counters
        print("grid soup n=%d %-10s %dx%d spp %d: %9.3f ms  %8.1f Mrays/s %8.1f Msamples/s  cells/ray %.2f tests/ray %.2f exact/ray %.2f" % (
            n, kernel, W, H, spp, best, c["rays"] / 1e3 / best, c["samples"] / 1e3 / best, c["cells_visited"] / c["rays"],
            c["tri_tests"] / c["rays"], c["tri_tests_executed"] / c["rays"]), flush=True)
