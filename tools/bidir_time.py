import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "scenes"))
import tempfile
import opencl_montecarlo_path_tracing_b200 as pt
import write_scenes
d = tempfile.mkdtemp()
write_scenes.write_variant("bidir", d)
sc = pt.load_scene_dir(d, "bidir")
with pt.Renderer(0) as r:
    r.set_scene(sc)
    lt = min(r.light_tracer((1, 2, 3, 4), 512) for _ in range(5))
    for (w, h) in ((512, 512), (1920, 1080)):
        for mem in ("smem", "const"):
            best = min(r.render("bidir", w, h, (1, 2, 3, 4), scene_mem=mem, read_image=False).ms for _ in range(6))
            print("bidir", w, h, mem, "render %.3f ms  light %.4f ms" % (best, lt), flush=True)
