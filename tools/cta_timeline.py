#!/usr/bin/env python3
"""Timeline of the big-grid megakernel's CTAs for one rank's share of the config-5 frame (PT_CTA_TIMES=1 diagnostics):
how many CTAs run at each moment, and how much of the launch is spent in the ramp-down tail."""
import os, sys
os.environ["PT_CTA_TIMES"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scenes"))
import ctypes as C
import numpy as np
import gen_mesh
import opencl_montecarlo_path_tracing_b200 as pt
tris = gen_mesh.soup(1 << 20); lo, hi = gen_mesh.bbox_like_reference(tris)
scene = pt.Scene(np.array([1024, 0, 0, 0, 145, 0, 0, 2048, 0], np.int32), np.array([4096, 0, 0, 0, 0, 0, 129, 0, 8192], np.int32),
                 tris, np.array([[10, 4, 10, 400], [15, 2, 7, 300]], np.float32), lo, hi)
W, H, spp, n = 3840, 2160, int(os.environ.get("RS_SPP", "64")), int(os.environ.get("RS_N", "8"))
with pt.Renderer(0) as r:
    r.set_scene(scene); r.build_grid(pt.grid_dims(scene))
    kw = dict(interleave=8, rank=0, nranks=n) if n > 1 else {}
    for it in range(4):
        res = r.render("grid", W, H, (1, 2, 3, 4), spp=spp, read_image=False, **kw)
    nrows = sum(min(8, H - s) for s in range(0, H, 8 * n)) if n > 1 else H
    nb = ((W + 15) // 16) * ((nrows + 7) // 8)
    buf = np.zeros(2 * nb, np.uint64)
    assert r._l.pt_debug_read_scratch(r.ctx, buf.ctypes.data_as(C.c_void_p), 256, buf.nbytes) == 0
    t = buf.reshape(nb, 2).astype(np.float64)
    t0 = t[:, 0].min(); start = (t[:, 0] - t0) / 1e6; end = (t[:, 1] - t0) / 1e6
    dur = end - start
    total = end.max()
    print("kernel %.3f ms (event), CTAs %d, span %.3f ms; CTA duration mean %.3f median %.3f max %.3f ms" % (res.ms, nb, total, dur.mean(), np.median(dur), dur.max()))
    print("sum of CTA durations / (span x resident slots %d) = %.3f" % (148 * 8, dur.sum() / (total * 148 * 8)))
    last_start = start.max()
    print("last CTA starts at %.3f ms (%.1f %% of the span); work after that point: %.1f %% of slot-time" % (
        last_start, 100 * last_start / total, 100 * np.clip(end - last_start, 0, None).sum() / ((total - last_start) * 148 * 8)))
    for q in range(10):
        a, b = total * q / 10, total * (q + 1) / 10
        busy = (np.minimum(end, b) - np.maximum(start, a)).clip(0).sum() / ((b - a) * 148 * 8)
        print("  decile %d: %.1f %% of the CTA slots busy, mean duration of CTAs started here %.3f ms" % (q, 100 * busy, dur[(start >= a) & (start < b)].mean() if ((start >= a) & (start < b)).any() else 0))
