import os, sys, tempfile
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scenes')
import write_scenes
import opencl_montecarlo_path_tracing_b200 as pt
with pt.Renderer(0) as r, tempfile.TemporaryDirectory() as tmp:
    for v in ("base", "lmem"):
        d = os.path.join(tmp, v); write_scenes.write_variant(v, d)
        r.set_scene(pt.load_scene_dir(d, v))
        for k, mem in (("mega", "smem"), ("mega", "const"), ("persistent", "const"), ("spec", "smem")):
            for cull in (True, False):
                res = r.render(v, 512, 512, (1, 2, 3, 4), kernel=k, scene_mem=mem, rows=(384, 392), read_image=False, cull=cull)
                res = r.render(v, 512, 512, (1, 2, 3, 4), kernel=k, scene_mem=mem, rows=(384, 392), read_image=False, cull=cull)
                print(v, k, mem, "cull" if cull else "nocull", "%.3f ms" % res.ms, res.counters["tri_tests_executed"], flush=True)
