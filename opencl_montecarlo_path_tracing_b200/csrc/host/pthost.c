/* pthost.c — see include/pthost.h.  C99, no CUDA. */
#define _GNU_SOURCE
#include "pthost.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
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

/* ---- triangles.txt --------------------------------------------------------------------------------------------------
 * The reference reads it with unchecked fgets() + atof(), 13 lines per triangle (3 x (x, y, z, blank) + blank), in a
 * `while (!feof(f) && n < MAX_TRIANGLES)` loop (CLSuperPathTracer.c:62-107).  A million triangles are 13.6 M lines, so
 * the file is read into memory once and the triangles whose 13 lines are all complete are converted on several threads
 * — with the SAME conversion, atof() on a private copy of each line — while the end of the file (where fgets() starts
 * returning short or no lines and the reference replays stale buffers) runs through a byte-exact emulation of the loop. */
struct mem_file { const char *p; size_t size, pos; int eof; };

/* fgets(buf, LINE_MAX_LEN, f) on the memory image, unchecked: buf keeps its content when nothing is left */
static void mem_next_line(struct mem_file *m, char *buf) {
    if (m->pos >= m->size) { m->eof = 1; return; }
    size_t n = 0;
    while (n < LINE_MAX_LEN - 1 && m->pos < m->size) {
        const char c = m->p[m->pos++];
        buf[n++] = c;
        if (c == '\n') break;
    }
    buf[n] = 0;
    if (n > 0 && buf[n - 1] != '\n' && n < LINE_MAX_LEN - 1) m->eof = 1;     /* ran into the end of the file while reading */
}

static float line_to_float(const char *line, size_t len) {
    char tmp[LINE_MAX_LEN];
    if (len > LINE_MAX_LEN - 1) len = LINE_MAX_LEN - 1;
    memcpy(tmp, line, len);
    tmp[len] = 0;
    return (float)atof(tmp);
}

struct tri_job {
    const char *buf;
    const size_t *start;      /* byte offset of the first line of every bulk triangle */
    size_t t0, t1;
    float *tris;
    float lo[3], hi[3];
};

static void *tri_worker(void *arg) {
    struct tri_job *j = (struct tri_job *)arg;
    for (int a = 0; a < 3; ++a) { j->lo[a] = FLT_MAX; j->hi[a] = FLT_MIN; }
    for (size_t n = j->t0; n < j->t1; ++n) {
        const char *p = j->buf + j->start[n];
        float *t = j->tris + 12 * n;
        for (int v = 0; v < 3; ++v) {
            for (int a = 0; a < 3; ++a) {
                const char *e = (const char *)strchr(p, '\n');          /* every line of a bulk triangle is complete */
                const float c = line_to_float(p, (size_t)(e - p) + 1);
                if (c < j->lo[a]) j->lo[a] = c;
                if (c > j->hi[a]) j->hi[a] = c;
                t[4 * v + a] = c;
                p = e + 1;
            }
            t[4 * v + 3] = 0.0f;
            p = strchr(p, '\n') + 1;                                     /* blank line closing the vertex */
        }
    }
    return NULL;
}

int pth_parse_triangles(const char *path, int max_triangles, float **out, float box_min[4], float box_max[4]) {
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    const long fsize = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *buf = (char *)malloc((size_t)(fsize > 0 ? fsize : 0) + 1);
    const size_t size = fsize > 0 ? fread(buf, 1, (size_t)fsize, f) : 0;
    fclose(f);
    buf[size] = 0;

    /* complete lines, and where every 13th starts; an over-long line (fgets would split it) or a NUL byte (string functions
     * would stop at it) sends the whole file through the emulation instead */
    size_t nlines = 0, ntri_lines = 0, cap_starts = 1024, nstarts = 0;
    size_t *start = (size_t *)malloc(cap_starts * sizeof(size_t));
    int plain = memchr(buf, 0, size) == NULL;
    for (size_t pos = 0; pos < size && plain;) {
        const char *e = (const char *)memchr(buf + pos, '\n', size - pos);
        if (!e) break;
        if ((size_t)(e - (buf + pos)) + 1 > LINE_MAX_LEN - 1) { plain = 0; break; }
        if (nlines % 13 == 0) {
            if (nstarts == cap_starts) { cap_starts *= 2; start = (size_t *)realloc(start, cap_starts * sizeof(size_t)); }
            start[nstarts++] = pos;
        }
        ++nlines;
        pos = (size_t)(e - buf) + 1;
    }
    ntri_lines = plain ? nlines / 13 : 0;
    size_t nbulk = ntri_lines < (size_t)(max_triangles > 0 ? max_triangles : 0) ? ntri_lines : (size_t)(max_triangles > 0 ? max_triangles : 0);
    if (nbulk > 0) --nbulk;          /* the last complete triangle goes through the emulation: it leaves the line buffers as the loop would */

    size_t cap = nbulk + 1024;
    float *tris = (float *)malloc(cap * 12 * sizeof(float));
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    float hi[3] = {FLT_MIN, FLT_MIN, FLT_MIN};   /* smallest positive float, as the reference */
    if (nbulk > 0) {
        long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
        int nth = (int)(ncpu < 1 ? 1 : (ncpu > 16 ? 16 : ncpu));
        if ((size_t)nth > nbulk / 4096 + 1) nth = (int)(nbulk / 4096 + 1);
        struct tri_job jobs[16];
        pthread_t th[16];
        int started[16] = {0};
        for (int k = 0; k < nth; ++k) {
            jobs[k].buf = buf; jobs[k].start = start; jobs[k].tris = tris;
            jobs[k].t0 = nbulk * (size_t)k / (size_t)nth; jobs[k].t1 = nbulk * (size_t)(k + 1) / (size_t)nth;
            if (k > 0) started[k] = pthread_create(&th[k], NULL, tri_worker, &jobs[k]) == 0;
        }
        tri_worker(&jobs[0]);
        for (int k = 1; k < nth; ++k) { if (started[k]) pthread_join(th[k], NULL); else tri_worker(&jobs[k]); }
        for (int k = 0; k < nth; ++k)
            for (int a = 0; a < 3; ++a) {
                if (jobs[k].lo[a] < lo[a]) lo[a] = jobs[k].lo[a];
                if (jobs[k].hi[a] > hi[a]) hi[a] = jobs[k].hi[a];
            }
    }

    /* the reference's loop, from triangle nbulk on */
    struct mem_file m = {buf, size, nbulk > 0 ? start[nbulk] : 0, 0};
    char coord[3][LINE_MAX_LEN] = {"", "", ""};
    size_t n = nbulk;
    while (!m.eof && n < (size_t)(max_triangles > 0 ? max_triangles : 0)) {
        if (n == cap) {
            cap *= 2;
            tris = (float *)realloc(tris, cap * 12 * sizeof(float));
        }
        float *t = tris + 12 * n;
        for (int v = 0; v < 3; ++v) {
            for (int a = 0; a < 3; ++a) mem_next_line(&m, coord[a]);
            for (int a = 0; a < 3; ++a) {
                float c = (float)atof(coord[a]);
                if (c < lo[a]) lo[a] = c;
                if (c > hi[a]) hi[a] = c;
                t[4 * v + a] = c;
            }
            t[4 * v + 3] = 0.0f;
            mem_next_line(&m, coord[0]);    /* blank line closing the vertex */
        }
        mem_next_line(&m, coord[0]);        /* blank line closing the triangle */
        ++n;
    }
    free(start);
    free(buf);
    for (int a = 0; a < 3; ++a) {
        if (box_min) box_min[a] = lo[a];
        if (box_max) box_max[a] = hi[a];
    }
    if (box_min) box_min[3] = 0.0f;
    if (box_max) box_max[3] = 0.0f;
    *out = tris;
    return (int)n;
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
