#!/usr/bin/env python3
"""Renders the headline frames once each with the AUTO kernels (for the final ncu captures of a round)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scenes"))
import numpy as np
import gen_mesh, write_scenes
import opencl_montecarlo_path_tracing_b200 as pt
with pt.Renderer(0) as r, tempfile.TemporaryDirectory() as tmp:
    for v in ("nodof", "base", "grid"):
        d = os.path.join(tmp, v); write_scenes.write_variant(v, d)
        scene = pt.load_scene_dir(d, v); r.set_scene(scene)
        if v == "grid":
            r.build_grid(pt.grid_dims(scene))
        for it in range(2):
            res = r.render(v, 512, 512, (1, 2, 3, 4), read_image=False)
        print(v, res.ms, flush=True)
    tris = gen_mesh.soup(1 << 20); lo, hi = gen_mesh.bbox_like_reference(tris)
    scene = pt.Scene(scene.spheres, scene.squares, tris, scene.lights, lo, hi)
    r.set_scene(scene); r.build_grid(pt.grid_dims(scene))
    for it in range(2):
        res = r.render("grid", 1920, 1080, (1, 2, 3, 4), spp=4, read_image=False)
    print("soup", res.ms, flush=True)
