#!/usr/bin/env python3
"""base / lmem default scenes (and torus / 1080p): PT_KERNEL_SPEC against the other flavours.  env QS_W QS_H QS_SPP"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scenes"))
import write_scenes
import opencl_montecarlo_path_tracing_b200 as pt
W, H, spp = int(os.environ.get("QS_W", "512")), int(os.environ.get("QS_H", "512")), int(os.environ.get("QS_SPP", "64"))
with pt.Renderer(0) as r, tempfile.TemporaryDirectory() as tmp:
    for v, mesh in (("base", None), ("lmem", None), ("base", "torus")):
        d = os.path.join(tmp, v + (mesh or "")); write_scenes.write_variant(v, d, mesh=mesh)
        r.set_scene(pt.load_scene_dir(d, v))
        for kernel in os.environ.get("QS_KERNELS", "auto,mega,persistent,spec").split(","):
            best = 1e9
            for it in range(5):
                res = r.render(v, W, H, (1, 2, 3, 4), spp=spp, kernel=kernel, read_image=False)
                best = min(best, res.ms)
            c = res.counters
            if kernel == "spec":
                import ctypes as C, numpy as np
                q = np.zeros(2, np.uint32)
                r._l.pt_debug_read_scratch(r.ctx, q.ctypes.data_as(C.c_void_p), 0, 8)
                print("      queued pixels: %d of %d (%.1f %%)" % (q[0], W * H, 100.0 * q[0] / (W * H)))
            print("%-5s %-6s %-10s %dx%dx%d: %8.3f ms  %9.1f Mrays/s  executed tri tests %d" % (v, mesh or "", kernel, W, H, spp, best, c["rays"] / 1e3 / best, c["tri_tests_executed"]), flush=True)
