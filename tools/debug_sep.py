import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scenes"))
import numpy as np
import opencl_montecarlo_path_tracing_b200 as pt
import write_scenes
from oracle.pyoracle import OracleLib
o = OracleLib(0)
with tempfile.TemporaryDirectory() as d, pt.Renderer(0) as r:
    write_scenes.write_variant("nodof", d)
    scene = pt.load_scene_dir(d, "nodof"); r.set_scene(scene)
    osc = o.load_scene_dir(d, "nodof")
    rows = (196, 260)
    ref = o.render("nodof", 512, 512, (1, 2, 3, 4), osc, rows=rows)
    for cull in (True, False):
        for kernel in ("mega", "persistent"):
            res = r.render("nodof", 512, 512, (1, 2, 3, 4), rows=rows, arith="separate", kernel=kernel, want_accum=True, want_rng=True, cull=cull)
            a = res.accum[rows[0]:rows[1]].view(np.uint32); b = ref["accum"][rows[0]:rows[1]].view(np.uint32)
            bad = np.argwhere((a != b).any(axis=2))
            print("cull", cull, kernel, "bad pixels", len(bad), [(int(y) + rows[0], int(x), res.accum[y + rows[0], x].tolist(), ref["accum"][y + rows[0], x].tolist()) for y, x in bad[:3]])
            rs = res.rng_state.reshape(8 * 512, 8 * 512, 4)[8 * rows[0]:8 * rows[1]]; rr = ref["rng_state"].reshape(8 * 512, 8 * 512, 4)[8 * rows[0]:8 * rows[1]]
            print("   rng differing samples:", int((rs != rr).any(axis=2).sum()), "counters", res.counters["rays"], ref["counters"]["rays"], res.counters["shadow_rays"], ref["counters"]["shadow_rays"])
