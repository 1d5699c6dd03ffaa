#!/usr/bin/env python3
"""config 3 frames (base scene with triangles.txt / torus.txt, 1920x1080, QC_SPP spp): kernel time per launch."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scenes"))
import write_scenes
import opencl_montecarlo_path_tracing_b200 as pt
spp = int(os.environ.get("QC_SPP", "1024"))
with pt.Renderer(0) as r, tempfile.TemporaryDirectory() as tmp:
    for mesh in (None, "torus"):
        d = os.path.join(tmp, mesh or "base"); write_scenes.write_variant("base", d, mesh=mesh)
        r.set_scene(pt.load_scene_dir(d, "base"))
        for it in range(4):
            res = r.render("base", 1920, 1080, (1, 2, 3, 4), spp=spp, read_image=False)
            print(mesh or "triangles", "launch %d: %.3f ms" % (it, res.ms), flush=True)
