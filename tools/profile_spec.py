#!/usr/bin/env python3
"""base 512x512x64 default scene with the AUTO kernels (PT_KERNEL_SPEC's two passes), 3 launches (ncu target)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scenes"))
import write_scenes
import opencl_montecarlo_path_tracing_b200 as pt
v = os.environ.get("PS_VARIANT", "base")
with pt.Renderer(0) as r, tempfile.TemporaryDirectory() as tmp:
    d = os.path.join(tmp, v); write_scenes.write_variant(v, d)
    r.set_scene(pt.load_scene_dir(d, v))
    for it in range(3):
        res = r.render(v, 512, 512, (1, 2, 3, 4), read_image=False)
    print(v, res.ms, res.counters)
