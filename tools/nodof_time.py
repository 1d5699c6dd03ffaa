#!/usr/bin/env python3
"""NoDoF kernel flavours side by side (best of 8 launches, device time)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scenes"))
import write_scenes
import opencl_montecarlo_path_tracing_b200 as pt
d = tempfile.mkdtemp(); write_scenes.write_variant("nodof", d)
with pt.Renderer(0) as r:
    r.set_scene(pt.load_scene_dir(d, "nodof"))
    for (w, h) in ((512, 512), (1920, 1080), (256, 256)):
        for k in ("mega", "persistent"):
            for mem in ("const", "smem"):
                best = min(r.render("nodof", w, h, (1, 2, 3, 4), kernel=k, scene_mem=mem, read_image=False).ms for _ in range(8))
                print("nodof %dx%d %-10s %-5s %.4f ms" % (w, h, k, mem, best), flush=True)
