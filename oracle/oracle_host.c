/*
 * oracle_host.c — restatement of the reference HOST-side helpers of the hot path:
 * scene parsers, camera basis, grid sizing/binning, PAM writer.
 * TEST INFRASTRUCTURE ONLY — see oracle.h.  Citations relative to /root/reference.
 */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXLINE 256   /* CLSuperPathTracer.c:13 */

/* ---- CLSuperPathTracer.c:26-50: float helpers (Normalize goes through a DOUBLE sqrt and the
 * quotient is narrowed to float when passed as ScalarTimesVector's `float scalar`) ---- */
typedef struct { float x, y, z; } H3;
static H3 h_scale(float s, H3 v) { H3 r = {s * v.x, s * v.y, s * v.z}; return r; }
static H3 h_add(H3 a, H3 b) { H3 r = {a.x + b.x, a.y + b.y, a.z + b.z}; return r; }
static float h_dot(H3 a, H3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static H3 h_cross(H3 a, H3 b) {
    H3 r = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    return r;
}
static H3 h_normalize(H3 v) { return h_scale((float)(1 / sqrt((double)h_dot(v, v))), v); }

/* CLSuperPathTracer.c:236-243 */
void oracle_camera(float cam_forward[4], float cam_up[4], float cam_right[4], float eye_offset[4]) {
    H3 z = {0, 0, -1};
    H3 fwd = {-6, -16, 0};
    fwd = h_normalize(fwd);
    H3 up = h_scale((float)0.002, h_normalize(h_cross(z, fwd)));
    H3 right = h_scale((float)0.002, h_normalize(h_cross(fwd, up)));
    H3 eye = h_add(h_scale((float)(-256), h_add(up, right)), fwd);
    cam_forward[0] = fwd.x; cam_forward[1] = fwd.y; cam_forward[2] = fwd.z; cam_forward[3] = 0;
    cam_up[0] = up.x; cam_up[1] = up.y; cam_up[2] = up.z; cam_up[3] = 0;
    cam_right[0] = right.x; cam_right[1] = right.y; cam_right[2] = right.z; cam_right[3] = 0;
    eye_offset[0] = eye.x; eye_offset[1] = eye.y; eye_offset[2] = eye.z; eye_offset[3] = 0;
}

/* CLSuperPathTracer.c:62-74: do { fgets; atoi; } while (!feof && n < 9) */
int oracle_parse_array(const char *path, int32_t arr[9]) {
    char str[MAXLINE] = {0};
    int n = 0;
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    do {
        if (!fgets(str, MAXLINE, f)) { /* buffer keeps its previous content, as in the reference */ }
        arr[n] = atoi(str);
        n++;
    } while (!feof(f) && n < 9);
    fclose(f);
    return n;
}

static void bb_update(float v, float *mn, float *mx) {
    if (v < *mn) *mn = v;
    if (v > *mx) *mx = v;
}

/* CLSuperPathTracer.c:77-118; bbox tracking as ..._trianglegrid/CLSuperPathTracer.c:136-209
 * (running max starts at FLT_MIN, the smallest POSITIVE float — a reference quirk). */
int oracle_parse_triangles(const char *path, float *tris12, int max_triangles, float box_min[4], float box_max[4]) {
    char x[MAXLINE] = {0}, y[MAXLINE] = {0}, z[MAXLINE] = {0};
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {FLT_MIN, FLT_MIN, FLT_MIN};
    int n = 0;
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    while (!feof(f) && n < max_triangles) {
        float *t = tris12 + 12 * (size_t)n;
        for (int v = 0; v < 3; ++v) {
            if (!fgets(x, MAXLINE, f)) {}
            if (!fgets(y, MAXLINE, f)) {}
            if (!fgets(z, MAXLINE, f)) {}
            float fx = (float)atof(x), fy = (float)atof(y), fz = (float)atof(z);
            bb_update(fx, &mn[0], &mx[0]);
            bb_update(fy, &mn[1], &mx[1]);
            bb_update(fz, &mn[2], &mx[2]);
            t[4 * v + 0] = fx; t[4 * v + 1] = fy; t[4 * v + 2] = fz; t[4 * v + 3] = 0.0f;
            if (!fgets(x, MAXLINE, f)) {}      /* vertex separator line, ignored */
        }
        if (!fgets(x, MAXLINE, f)) {}          /* triangle separator line, ignored */
        n++;
    }
    fclose(f);
    if (box_min) { box_min[0] = mn[0]; box_min[1] = mn[1]; box_min[2] = mn[2]; box_min[3] = 0; }
    if (box_max) { box_max[0] = mx[0]; box_max[1] = mx[1]; box_max[2] = mx[2]; box_max[3] = 0; }
    return n;
}

/* CLSuperPathTracer.c:121-139 */
int oracle_parse_lights(const char *path, float lights[5][4]) {
    char b[4][MAXLINE] = {{0}};
    int n = 0;
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    while (!feof(f) && n < 5) {
        for (int k = 0; k < 4; ++k)
            if (!fgets(b[k], MAXLINE, f)) {}
        for (int k = 0; k < 4; ++k) lights[n][k] = (float)atof(b[k]);
        n++;
    }
    fclose(f);
    return n;
}

/* ..._trianglegrid/CLSuperPathTracer.c:476-484 */
void oracle_grid_dims(const float box_min[4], const float box_max[4], int ntriangles, float cell_size_modifier,
                      int32_t grid_res[4], float cell_size[4]) {
    float gs[3];
    for (int a = 0; a < 3; ++a) gs[a] = box_max[a] - box_min[a];
    float cube_root = (float)cbrt((double)(cell_size_modifier * ntriangles / (gs[0] * gs[1] * gs[2])));
    for (int a = 0; a < 3; ++a) {
        int r = (int)(floor((double)(gs[a] * cube_root)));
        r = r < 128 ? r : 128;
        r = r > 1 ? r : 1;
        grid_res[a] = r;
        cell_size[a] = gs[a] / r;
    }
    grid_res[3] = 0;
    cell_size[3] = 0;
}

static inline float cl_fmin(float x, float y) { if (x != x) return y; if (y != y) return x; return y < x ? y : x; }
static inline float cl_fmax(float x, float y) { if (x != x) return y; if (y != y) return x; return x < y ? y : x; }
static inline int f2i_rz_sat(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (-2147483647 - 1);
    return (int)f;
}
static inline int clampi(int v, int lo, int hi) { v = v < lo ? lo : v; return v > hi ? hi : v; }

/* Cell range of one triangle: ..._trianglegrid/pathtracer.ocl:314-321 */
static void tri_cell_range(const float *t, const float box_min[4], const int32_t res[4], const float cell[4],
                           int lo[3], int hi[3]) {
    for (int a = 0; a < 3; ++a) {
        float mn = cl_fmin(t[a], cl_fmin(t[4 + a], t[8 + a]));
        float mx = cl_fmax(t[a], cl_fmax(t[4 + a], t[8 + a]));
        lo[a] = clampi(f2i_rz_sat((mn - box_min[a]) / cell[a]), 0, res[a] - 1);
        hi[a] = clampi(f2i_rz_sat((mx - box_min[a]) / cell[a]), 0, res[a] - 1);
    }
}

uint64_t oracle_build_grid(const float *tris12, int ntris, const float box_min[4], const int32_t grid_res[4],
                           const float cell_size[4], int cap, uint32_t *cell_start, uint32_t *cell_refs) {
    size_t ncells = (size_t)grid_res[0] * grid_res[1] * grid_res[2];
    uint32_t *count = (uint32_t *)calloc(ncells + 1, sizeof(uint32_t));
    int lo[3], hi[3];
    for (int i = 0; i < ntris; ++i) {
        tri_cell_range(tris12 + 12 * (size_t)i, box_min, grid_res, cell_size, lo, hi);
        for (int z = lo[2]; z <= hi[2]; ++z)
            for (int y = lo[1]; y <= hi[1]; ++y)
                for (int x = lo[0]; x <= hi[0]; ++x) {
                    size_t c = (size_t)z * grid_res[0] * grid_res[1] + (size_t)y * grid_res[0] + x;
                    if ((int)count[c] < cap) count[c]++;
                }
    }
    uint64_t total = 0;
    for (size_t c = 0; c < ncells; ++c) { cell_start[c] = (uint32_t)total; total += count[c]; }
    cell_start[ncells] = (uint32_t)total;
    if (cell_refs) {
        memset(count, 0, (ncells + 1) * sizeof(uint32_t));
        for (int i = 0; i < ntris; ++i) {
            tri_cell_range(tris12 + 12 * (size_t)i, box_min, grid_res, cell_size, lo, hi);
            for (int z = lo[2]; z <= hi[2]; ++z)
                for (int y = lo[1]; y <= hi[1]; ++y)
                    for (int x = lo[0]; x <= hi[0]; ++x) {
                        size_t c = (size_t)z * grid_res[0] * grid_res[1] + (size_t)y * grid_res[0] + x;
                        if ((int)count[c] < cap) cell_refs[cell_start[c] + count[c]++] = (uint32_t)i;
                    }
        }
    }
    free(count);
    return total;
}

/* ---- VLP bounding box and VLP grid of CLSuperMetropolisPathTracer_vlpgrid --------------------------------------------
 * kernels reduceMinAndMax_lmem / reduceMinAndMax_lmem_nwg (metropolispathtracer.ocl:538-619): a light (x y z intensity)
 * with intensity 0 is a dummy (box FLT_MAX / FLT_MIN), every other one spans position -+ 16 * sqrt(intensity); the two
 * tree passes keep, per component, the smaller minimum / larger maximum (isless / isgreater + select: a NaN arriving
 * from the partner never wins; restated as "NaN components are skipped").  vmin / vmax as the host reads the Box back
 * (CLSuperMetropolisPathTracer.c:606-611). */
static int vlp_reach(const float *v, float lo[3], float hi[3]) {
    if (v[3] == 0.0f) return 0;
    volatile float r = 16.0f * sqrtf(v[3]);
    for (int a = 0; a < 3; ++a) {
        volatile float l = v[a] - r, h = v[a] + r;
        lo[a] = l; hi[a] = h;
    }
    return 1;
}

void oracle_vlp_bounds(const float *vpl4, int n, float vmin[4], float vmax[4]) {
    for (int a = 0; a < 3; ++a) { vmin[a] = FLT_MAX; vmax[a] = FLT_MIN; }
    vmin[3] = vmax[3] = 0.0f;
    for (int i = 0; i < n; ++i) {
        float lo[3], hi[3];
        if (!vlp_reach(vpl4 + 4 * (size_t)i, lo, hi)) continue;
        for (int a = 0; a < 3; ++a) {
            if (lo[a] < vmin[a]) vmin[a] = lo[a];
            if (hi[a] > vmax[a]) vmax[a] = hi[a];
        }
    }
}

/* kernel initVLPsGrid (metropolispathtracer.ocl:621-647): cell range = clamp(convert_int4(((pos -+ r) - boxmin) / cell_size),
 * 0, res - 1); every overlapped cell gets the light's index, at most `cap` (62) per cell.  The reference appends with
 * atomic_inc (any order); here ascending index order, what a serial run of the kernel produces. */
static void vlp_cell_range(const float *v, const float box_min[4], const int32_t res[4], const float cell[4], int lo[3], int hi[3], int *live) {
    float l[3], h[3];
    *live = vlp_reach(v, l, h);
    if (!*live) return;
    for (int a = 0; a < 3; ++a) {
        volatile float dl = l[a] - box_min[a], dh = h[a] - box_min[a];
        volatile float ql = dl / cell[a], qh = dh / cell[a];
        int cl = f2i_rz_sat(ql), ch = f2i_rz_sat(qh);
        lo[a] = cl < 0 ? 0 : (cl > res[a] - 1 ? res[a] - 1 : cl);
        hi[a] = ch < 0 ? 0 : (ch > res[a] - 1 ? res[a] - 1 : ch);
    }
}

uint64_t oracle_build_vlp_grid(const float *vpl4, int n, const float box_min[4], const int32_t grid_res[4], const float cell_size[4],
                               int cap, uint32_t *cell_start, uint32_t *cell_refs) {
    size_t ncells = (size_t)grid_res[0] * grid_res[1] * grid_res[2];
    uint32_t *count = (uint32_t *)calloc(ncells + 1, sizeof(uint32_t));
    int lo[3], hi[3], live;
    for (int pass = 0; pass < (cell_refs ? 2 : 1); ++pass) {
        if (pass == 1) memset(count, 0, (ncells + 1) * sizeof(uint32_t));
        for (int i = 0; i < n; ++i) {
            vlp_cell_range(vpl4 + 4 * (size_t)i, box_min, grid_res, cell_size, lo, hi, &live);
            if (!live) continue;
            for (int z = lo[2]; z <= hi[2]; ++z)
                for (int y = lo[1]; y <= hi[1]; ++y)
                    for (int x = lo[0]; x <= hi[0]; ++x) {
                        size_t c = (size_t)z * grid_res[0] * grid_res[1] + (size_t)y * grid_res[0] + x;
                        if ((int)count[c] < cap) {
                            if (pass == 1) cell_refs[cell_start[c] + count[c]] = (uint32_t)i;
                            count[c]++;
                        }
                    }
        }
        if (pass == 0) {
            uint64_t total = 0;
            for (size_t c = 0; c < ncells; ++c) { cell_start[c] = (uint32_t)total; total += count[c]; }
            cell_start[ncells] = (uint32_t)total;
        }
    }
    uint64_t total = cell_start[ncells];
    free(count);
    return total;
}

/* pamalign.h:212-238: "P7" header + raw RGBA bytes */
int oracle_save_pam(const char *path, int width, int height, const uint8_t *rgba8) {
    FILE *f = fopen(path, "wb");
    if (!f) return 1;
    fprintf(f, "P7\nWIDTH %u\nHEIGHT %u\nDEPTH %u\nMAXVAL %u\nTUPLTYPE %s\nENDHDR\n", (unsigned)width, (unsigned)height,
            4u, 255u, "RGB_ALPHA");
    fwrite(rgba8, 1, (size_t)width * height * 4, f);
    fclose(f);
    return 0;
}
