#!/usr/bin/env python3
"""Renders each default scene once per requested kernel kind (for ncu captures).
usage: tools/profile_scenes.py variant[:kernel[:mem]] ..."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scenes"))
import opencl_montecarlo_path_tracing_b200 as pt  # noqa: E402
import write_scenes  # noqa: E402

with pt.Renderer(0) as r, tempfile.TemporaryDirectory() as tmp:
    for spec in sys.argv[1:]:
        parts = spec.split(":")
        v = parts[0]; kernel = parts[1] if len(parts) > 1 else "auto"; mem = parts[2] if len(parts) > 2 else "auto"
        d = os.path.join(tmp, v)
        write_scenes.write_variant(v, d)
        scene = pt.load_scene_dir(d, v)
        r.set_scene(scene)
        if v == "grid":
            r.build_grid(pt.grid_dims(scene))
        for it in range(3):
            res = r.render(v, 512, 512, (1, 2, 3, 4), kernel=kernel, scene_mem=mem, read_image=False)
        print(spec, "%.3f ms" % res.ms, res.counters, flush=True)
