#!/usr/bin/env python3
"""Renders each headline frame twice with the AUTO kernels (for the ncu captures of a round; profile the 2nd launch)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scenes"))
import numpy as np
import gen_mesh, write_scenes
import opencl_montecarlo_path_tracing_b200 as pt
which = set((os.environ.get("PF_WHICH") or "nodof,base,lmem,grid,bidir,base1080,soup").split(","))
with pt.Renderer(0) as r, tempfile.TemporaryDirectory() as tmp:
    for v in ("nodof", "base", "lmem", "grid", "bidir"):
        if v not in which and not (v == "base" and "base1080" in which):
            continue
        d = os.path.join(tmp, v); write_scenes.write_variant(v, d)
        scene = pt.load_scene_dir(d, v); r.set_scene(scene)
        if v == "grid":
            r.build_grid(pt.grid_dims(scene))
        if v == "bidir":
            r.light_tracer((1, 2, 3, 4), 512)
        if v in which:
            for it in range(2):
                res = r.render(v, 512, 512, (1, 2, 3, 4), read_image=False)
            print(v, res.ms, flush=True)
        if v == "base" and "base1080" in which:
            for it in range(2):
                res = r.render(v, 1920, 1080, (1, 2, 3, 4), spp=32, read_image=False)
            print("base1080 spp32", res.ms, flush=True)
    if "soup" in which:
        d = os.path.join(tmp, "g"); write_scenes.write_variant("grid", d)
        scene = pt.load_scene_dir(d, "grid")
        tris = gen_mesh.soup(1 << 20); lo, hi = gen_mesh.bbox_like_reference(tris)
        scene = pt.Scene(scene.spheres, scene.squares, tris, scene.lights, lo, hi)
        r.set_scene(scene); r.build_grid(pt.grid_dims(scene))
        for it in range(2):
            res = r.render("grid", 1920, 1080, (1, 2, 3, 4), spp=4, read_image=False)
        print("soup", res.ms, flush=True)
