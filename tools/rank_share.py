#!/usr/bin/env python3
"""One rank's share of the strong-scaled config-5 frame (3840x2160, 1 M-triangle soup, RS_SPP spp) on ONE GPU:
8-row stripes of rank RS_RANK of RS_N, timed with the 4-warp and the 1-warp CTA megakernel (PT_MEGA_WARPS)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scenes"))
import numpy as np
import gen_mesh
import opencl_montecarlo_path_tracing_b200 as pt
tris = gen_mesh.soup(1 << 20); lo, hi = gen_mesh.bbox_like_reference(tris)
scene = pt.Scene(np.array([1024, 0, 0, 0, 145, 0, 0, 2048, 0], np.int32), np.array([4096, 0, 0, 0, 0, 0, 129, 0, 8192], np.int32),
                 tris, np.array([[10, 4, 10, 400], [15, 2, 7, 300]], np.float32), lo, hi)
W, H, spp = int(os.environ.get("RS_W", "3840")), int(os.environ.get("RS_H", "2160")), int(os.environ.get("RS_SPP", "64"))
with pt.Renderer(0) as r:
    r.set_scene(scene); r.build_grid(pt.grid_dims(scene))
    for n in [int(x) for x in os.environ.get("RS_N", "1,2,4,8").split(",")]:
        best = 1e9
        for it in range(int(os.environ.get('RS_REPS', '4'))):
            kw = dict(interleave=int(os.environ.get("RS_STRIPE", "8")), rank=int(os.environ.get("RS_RANK", "0")), nranks=n) if n > 1 else {}
            res = r.render("grid", W, H, (1, 2, 3, 4), spp=spp, kernel=os.environ.get("RS_KERNEL", "auto"), read_image=False, **kw)
            best = min(best, res.ms); print('   launch %d: %.3f ms' % (it, res.ms), flush=True)
        print("N=%d rank-0 share: %.3f ms  (x N = %.3f)  rays %d" % (n, best, best * n, res.counters["rays"]), flush=True)
