/*
 * pthost.h — host-side helpers of the drop-in CLSuperPathTracer programs: the scene-file readers,
 * camera basis, grid sizing and PAM writer that each reference main() carries inline.  Plain C, no
 * CUDA: compiled with gcc into the CLI programs and into libpthost.so (bound by tests via ctypes).
 *
 *   reference                                                         here
 *   parseArrayFromFile      CLSuperPathTracer.c:62-74                 pth_parse_bitmap
 *   parseTrianglesFromFile  CLSuperPathTracer.c:77-118 (+bbox: ..._trianglegrid/CLSuperPathTracer.c:136-209)
 *                                                                     pth_parse_triangles
 *   parseLightsFromFile     CLSuperPathTracer.c:121-139               pth_parse_lights
 *   camera set-up           CLSuperPathTracer.c:236-243               pth_camera
 *   grid sizing             ..._trianglegrid/CLSuperPathTracer.c:476-483   pth_grid_dims
 *   save_pam                pamalign.h:212-238                        pth_save_pam
 *   seeds                   CLSuperPathTracer.c:209                   pth_seeds
 *   load_pam (+ imgInfo)    pamalign.h:13-21, 52-131, 166-210         pth_load_pam / pth_image
 * Formats the reference lacks: pth_save_ppm (real P6), pth_save_png, pth_import_obj.
 */
#ifndef PTHOST_H
#define PTHOST_H
#include <stddef.h>
#include <stdint.h>
#include "ptcuda.h"
#ifdef __cplusplus
extern "C" {
#endif

/* 9-line bitmap file; returns lines consumed, -1 if the file cannot be opened (the reference would crash) */
int pth_parse_bitmap(const char *path, int32_t rows[9]);
/* Reads at most max_triangles triangles into a malloc'ed n x 12 float array (*out, caller frees).
 * box_min/box_max (may be NULL) receive the reference's running bounds (max starts at FLT_MIN). */
int pth_parse_triangles(const char *path, int max_triangles, float **out, float box_min[4], float box_max[4]);
/* print_lights != 0 echoes "Light %d: ..." like the lmem/NoDoF/grid hosts */
int pth_parse_lights(const char *path, float lights[5][4], int print_lights);
void pth_camera(pt_camera *cam);
void pth_grid_dims(const float box_min[4], const float box_max[4], int ntriangles, float cell_size_modifier, pt_grid *grid);
int pth_save_pam(const char *path, int width, int height, const void *rgba8);
/* imgInfo of pamalign.h:13-21 */
typedef struct pth_image {
    uint32_t width, height, channels, maxval;
    uint32_t depth;      /* bits per value: 8 or 16 */
    size_t data_size;
    void *data;          /* malloc'ed; 3-channel images are padded to 4 values per pixel, 16-bit values host-endian */
} pth_image;
/* load_pam (pamalign.h:166-210): returns 0 on success, 1 on any error (message on stderr, like the reference). */
int pth_load_pam(const char *path, pth_image *img);
void pth_free_image(pth_image *img);
/* Binary P6 PPM (alpha dropped) / 8-bit RGBA PNG of an RGBA8 frame such as pt_map_render returns. 0 = ok. */
int pth_save_ppm(const char *path, int width, int height, const void *rgba8);
int pth_save_png(const char *path, int width, int height, const void *rgba8);
/* Wavefront OBJ -> triangles.txt (the text layout parseTrianglesFromFile reads); coordinates scale*c + translate.
 * Returns the triangle count, -1 on error. */
long pth_import_obj(const char *obj_path, const char *triangles_txt_path, float scale, const float translate[3]);
/* PT_SEEDS=a,b,c,d if set, else the reference's wall-clock recipe masked to 27 bits */
void pth_seeds(uint32_t seeds[4]);

/* Complete drop-in program: argv, stdout lines, scene files from CWD and result.ppm as the reference
 * main() of the given PT_VARIANT_* (used by the four CLSuperPathTracer executables). */
int pth_cli_main(int variant, int argc, char **argv);
/* CLSuperMetropolisPathTracer_vlpgrid/CLSuperMetropolisPathTracer.c:429-720 (its kernels in FIX mode, ptcuda.h) */
int pth_cli_metropolis_main(int argc, char **argv);

#ifdef __cplusplus
}
#endif
#endif
