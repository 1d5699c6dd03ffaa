/* pthost.c — see include/pthost.h.  C99, no CUDA. */
#define _GNU_SOURCE
#include "pthost.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

enum { LINE_MAX_LEN = 256 };  /* the reference's #define MAX 256 */

/* One text line into buf; when the file is exhausted buf keeps what it held, exactly like an
 * unchecked fgets() — that is what makes a trailing newline replay the last record. */
static void next_line(FILE *f, char *buf) {
    char *r = fgets(buf, LINE_MAX_LEN, f);
    (void)r;
}

int pth_parse_bitmap(const char *path, int32_t rows[9]) {
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char line[LINE_MAX_LEN] = "";
    int n = 0;
    do {                                    /* do/while: at least one line is always consumed */
        next_line(f, line);
        rows[n++] = atoi(line);
    } while (!feof(f) && n < 9);
    fclose(f);
    return n;
}

int pth_parse_triangles(const char *path, int max_triangles, float **out, float box_min[4], float box_max[4]) {
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    size_t cap = 1024;
    float *tris = (float *)malloc(cap * 12 * sizeof(float));
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    float hi[3] = {FLT_MIN, FLT_MIN, FLT_MIN};   /* smallest positive float, as the reference */
    char coord[3][LINE_MAX_LEN] = {"", "", ""};
    int n = 0;
    while (!feof(f) && n < max_triangles) {
        if ((size_t)n == cap) {
            cap *= 2;
            tris = (float *)realloc(tris, cap * 12 * sizeof(float));
        }
        float *t = tris + 12 * (size_t)n;
        for (int v = 0; v < 3; ++v) {
            for (int a = 0; a < 3; ++a) next_line(f, coord[a]);
            for (int a = 0; a < 3; ++a) {
                float c = (float)atof(coord[a]);
                if (c < lo[a]) lo[a] = c;
                if (c > hi[a]) hi[a] = c;
                t[4 * v + a] = c;
            }
            t[4 * v + 3] = 0.0f;
            next_line(f, coord[0]);         /* blank line closing the vertex */
        }
        next_line(f, coord[0]);             /* blank line closing the triangle */
        ++n;
    }
    fclose(f);
    for (int a = 0; a < 3; ++a) {
        if (box_min) box_min[a] = lo[a];
        if (box_max) box_max[a] = hi[a];
    }
    if (box_min) box_min[3] = 0.0f;
    if (box_max) box_max[3] = 0.0f;
    *out = tris;
    return n;
}

int pth_parse_lights(const char *path, float lights[5][4], int print_lights) {
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char field[4][LINE_MAX_LEN] = {"", "", "", ""};
    int n = 0;
    while (!feof(f) && n < 5) {
        for (int k = 0; k < 4; ++k) next_line(f, field[k]);
        for (int k = 0; k < 4; ++k) lights[n][k] = (float)atof(field[k]);
        if (print_lights)
            printf("Light %d: %f %f %f %f\n", n, lights[n][0], lights[n][1], lights[n][2], lights[n][3]);
        ++n;
    }
    fclose(f);
    return n;
}

/* The reference normalises through a double-precision sqrt and narrows the reciprocal to float before
 * scaling (Normalize -> ScalarTimesVector(float scalar, ...), CLSuperPathTracer.c:31-50). */
static void unit3(const float v[3], float out[3]) {
    float len2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    float s = (float)(1 / sqrt((double)len2));
    for (int a = 0; a < 3; ++a) out[a] = s * v[a];
}
static void cross3f(const float a[3], const float b[3], float out[3]) {
    out[0] = a[1] * b[2] - a[2] * b[1];
    out[1] = a[2] * b[0] - a[0] * b[2];
    out[2] = a[0] * b[1] - a[1] * b[0];
}

void pth_camera(pt_camera *cam) {
    const float minus_z[3] = {0, 0, -1};
    const float look[3] = {-6, -16, 0};
    const float k = (float)0.002;
    float fwd[3], tmp[3], up[3], right[3];
    unit3(look, fwd);
    cross3f(minus_z, fwd, tmp);
    unit3(tmp, up);
    for (int a = 0; a < 3; ++a) up[a] = k * up[a];
    cross3f(fwd, up, tmp);
    unit3(tmp, right);
    for (int a = 0; a < 3; ++a) right[a] = k * right[a];
    for (int a = 0; a < 3; ++a) {
        cam->cam_forward[a] = fwd[a];
        cam->cam_up[a] = up[a];
        cam->cam_right[a] = right[a];
        cam->eye_offset[a] = (float)(-256) * (up[a] + right[a]) + fwd[a];  /* -256 regardless of the image size */
    }
    cam->cam_forward[3] = cam->cam_up[3] = cam->cam_right[3] = cam->eye_offset[3] = 0.0f;
}

void pth_grid_dims(const float box_min[4], const float box_max[4], int ntriangles, float modifier, pt_grid *g) {
    float size[3];
    for (int a = 0; a < 3; ++a) {
        g->box_min[a] = box_min[a];
        g->box_max[a] = box_max[a];
        size[a] = box_max[a] - box_min[a];
    }
    g->box_min[3] = g->box_max[3] = 0.0f;
    float density = modifier * ntriangles / (size[0] * size[1] * size[2]);
    float root = (float)cbrt((double)density);
    for (int a = 0; a < 3; ++a) {
        int r = (int)floor((double)(size[a] * root));
        if (r > 128) r = 128;
        if (r < 1) r = 1;
        g->res[a] = r;
        g->cell_size[a] = size[a] / r;
    }
    g->res[3] = 0;
    g->cell_size[3] = 0.0f;
    g->max_refs_per_cell = 62;
}

int pth_save_pam(const char *path, int width, int height, const void *rgba8) {
    FILE *f = fopen(path, "wb");
    if (!f) {
        fprintf(stderr, "could not open %s for writing\n", path);
        return 1;
    }
    fputs("P7\n", f);
    fprintf(f, "WIDTH %u\nHEIGHT %u\nDEPTH %u\nMAXVAL %u\nTUPLTYPE %s\nENDHDR\n", (unsigned)width, (unsigned)height, 4u, 255u,
            "RGB_ALPHA");
    size_t n = (size_t)width * height * 4;
    int bad = fwrite(rgba8, 1, n, f) != n;
    fclose(f);
    return bad;
}

static uint64_t cycle_counter(void) {
#if defined(__x86_64__) || defined(__i386__)
    uint32_t lo, hi;
    __asm__ __volatile__("rdtsc" : "=a"(lo), "=d"(hi));
    return ((uint64_t)hi << 32) | lo;
#else
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
#endif
}

void pth_seeds(uint32_t seeds[4]) {
    const char *env = getenv("PT_SEEDS");
    unsigned long v[4];
    if (env && sscanf(env, "%lu,%lu,%lu,%lu", &v[0], &v[1], &v[2], &v[3]) == 4) {
        for (int i = 0; i < 4; ++i) seeds[i] = (uint32_t)v[i];
        return;
    }
    const uint32_t mask = 134217727u;   /* 27 bits */
    int pid = (int)getpid();
    clock_t ck = clock();
    seeds[0] = (uint32_t)time(0) & mask;
    seeds[1] = (uint32_t)(pid * pid * pid) & mask;
    seeds[2] = (uint32_t)(ck * ck) & mask;
    seeds[3] = (uint32_t)cycle_counter() & mask;
}
