#!/usr/bin/env python3
"""Where the end-to-end call (pt_render_host) spends its time on the NoDoF benchmark frame."""
import ctypes as C, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scenes"))
import numpy as np
import write_scenes
import opencl_montecarlo_path_tracing_b200 as pt
from opencl_montecarlo_path_tracing_b200 import _lib
lib = _lib.cuda_lib()
d = tempfile.mkdtemp(); write_scenes.write_variant("nodof", d)
scene = pt.load_scene_dir(d, "nodof")
r = pt.Renderer(0); cs = scene.to_c()
p = pt.make_params("nodof", 512, 512, (1, 2, 3, 4))
img = np.zeros((512, 512, 4), np.uint8)
N = 300
def t(fn, n=N):
    for _ in range(20): fn()
    r.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    r.synchronize(); return (time.perf_counter() - t0) * 1e6 / n
def full():
    assert lib.pt_render_host(r.ctx, C.byref(cs), None, C.byref(r.cam), C.byref(p), img.ctypes.data_as(C.POINTER(C.c_uint8))) == 0
def setscene():
    lib.pt_set_scene(r.ctx, C.byref(cs))
def launch():
    e = lib.pt_launch_pathtracer(r.ctx, C.byref(r.cam), C.byref(p)); lib.pt_wait(e); lib.pt_release_event(e)
def launch_map():
    e = lib.pt_launch_pathtracer(r.ctx, C.byref(r.cam), C.byref(p)); lib.pt_map_render(r.ctx, None); lib.pt_release_event(e)
src = np.zeros((512, 512, 4), np.uint8)
def memcpy():
    img[...] = src
print("pt_render_host        %.1f us" % t(full))
print("pt_set_scene          %.1f us" % t(setscene))
print("launch + wait         %.1f us" % t(launch))
print("launch + map_render   %.1f us" % t(launch_map))
print("host memcpy 1 MB      %.1f us" % t(memcpy))
