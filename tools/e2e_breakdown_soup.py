#!/usr/bin/env python3
"""Host-side cost of the end-to-end call on the 1 M-triangle soup (config 4/5 scene): scene upload and grid build."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scenes"))
import numpy as np
import gen_mesh
import opencl_montecarlo_path_tracing_b200 as pt
from opencl_montecarlo_path_tracing_b200 import _lib
lib = _lib.cuda_lib()
tris = gen_mesh.soup(1 << 20); lo, hi = gen_mesh.bbox_like_reference(tris)
scene = pt.Scene(np.array([1024, 0, 0, 0, 145, 0, 0, 2048, 0], np.int32), np.array([4096, 0, 0, 0, 0, 0, 129, 0, 8192], np.int32),
                 tris, np.array([[10, 4, 10, 400], [15, 2, 7, 300]], np.float32), lo, hi)
r = pt.Renderer(0); cs = scene.to_c(); g = pt.grid_dims(scene)
def t(fn, n=10):
    for _ in range(2): fn()
    r.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    r.synchronize(); return (time.perf_counter() - t0) * 1e3 / n
def setscene(): lib.pt_set_scene(r.ctx, C.byref(cs))
def build():
    e = lib.pt_build_grid(r.ctx, C.byref(g)); lib.pt_wait(e); lib.pt_release_event(e)
print("pt_set_scene  %.2f ms" % t(setscene))
print("pt_build_grid %.2f ms" % t(build))
