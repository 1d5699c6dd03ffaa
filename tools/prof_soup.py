#!/usr/bin/env python3
"""Renders the 1 M-triangle soup frame (1920x1080) PS_REPS times with kernel PS_KERNEL at PS_SPP spp (ncu target)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scenes"))
import numpy as np
import gen_mesh
import opencl_montecarlo_path_tracing_b200 as pt
tris = gen_mesh.soup(1 << 20); lo, hi = gen_mesh.bbox_like_reference(tris)
scene = pt.Scene(np.array([1024, 0, 0, 0, 145, 0, 0, 2048, 0], np.int32), np.array([4096, 0, 0, 0, 0, 0, 129, 0, 8192], np.int32),
                 tris, np.array([[10, 4, 10, 400], [15, 2, 7, 300]], np.float32), lo, hi)
W = int(os.environ.get("PS_W", "1920")); H = int(os.environ.get("PS_H", "1080"))
with pt.Renderer(0) as r:
    r.set_scene(scene); r.build_grid(pt.grid_dims(scene))
    for k in os.environ.get("PS_KERNEL", "auto").split(","):
        for it in range(int(os.environ.get("PS_REPS", "2"))):
            res = r.render("grid", W, H, (1, 2, 3, 4), spp=int(os.environ.get("PS_SPP", "4")), kernel=k, read_image=False, dead_rays=os.environ.get("PS_DEAD", "auto"))
        print(k, res.ms, res.counters, flush=True)
