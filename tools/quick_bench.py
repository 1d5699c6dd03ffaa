#!/usr/bin/env python3
"""Quick device-time comparison of kernel kinds / scene memories on the default scenes (GPU box).
usage: tools/quick_bench.py [variant ...]   prints one line per configuration."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scenes"))
import opencl_montecarlo_path_tracing_b200 as pt  # noqa: E402
import write_scenes  # noqa: E402

variants = sys.argv[1:] or ["nodof", "base", "lmem", "grid"]
W = int(os.environ.get("QB_W", "512")); H = int(os.environ.get("QB_H", "512")); SPP = int(os.environ.get("QB_SPP", "64"))
with pt.Renderer(0) as r, tempfile.TemporaryDirectory() as tmp:
    for v in variants:
        d = os.path.join(tmp, v)
        write_scenes.write_variant(v, d)
        scene = pt.load_scene_dir(d, v)
        r.set_scene(scene)
        if v == "grid":
            r.build_grid(pt.grid_dims(scene))
        for kernel in os.environ.get("QB_KERNELS", "mega,persistent").split(","):
            for mem in (("const",) if kernel == "wavefront" else ("const", "smem")):
                spp = 64 if v == "nodof" else SPP
                best = 1e9
                for it in range(5):
                    res = r.render(v, W, H, (1, 2, 3, 4), spp=spp, kernel=kernel, scene_mem=mem, read_image=False, cull=os.environ.get("QB_CULL", "1") == "1")
                    best = min(best, res.ms)
                c = res.counters
                print("%-6s %-10s %-5s %4dx%-4d spp %-4d  %8.3f ms  %9.1f Mrays/s  %8.1f Msamples/s  rays %d tri_exec %d" % (
                    v, kernel, mem, W, H, spp, best, c["rays"] / 1e3 / best, c["samples"] / 1e3 / best, c["rays"], c["tri_tests_executed"]), flush=True)
