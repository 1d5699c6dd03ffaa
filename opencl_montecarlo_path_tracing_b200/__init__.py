"""B200-native drop-in for the CLSuperPathTracer hot path of
JustAToaster/OpenCL_MonteCarlo_Path_Tracing.

Layout:  csrc/ (CUDA kernels + C ABI + C host programs)   lib/ bin/ (built in-tree)
         _lib.py (ctypes binding)   api.py (Scene / Renderer convenience layer)
The drop-in boundary is the C ABI in include/ptcuda.h; see INTEGRATION.md.
"""
from .api import (PtError, Renderer, RenderResult, Scene, camera, default_max_triangles, grid_dims, load_scene_dir,  # noqa: F401
                  make_params, save_pam, load_pam, save_ppm, save_png, import_obj, vlp_grid_dims)

__all__ = ["PtError", "Renderer", "RenderResult", "Scene", "camera", "default_max_triangles", "grid_dims",
           "load_scene_dir", "make_params", "save_pam", "load_pam", "save_ppm", "save_png", "import_obj", "vlp_grid_dims"]
