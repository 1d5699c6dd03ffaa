/*
 * oracle_cli.c — command-line front end of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 *   oracle_cli <base|lmem|nodof|grid|bidir> [img_width] [img_height] [CELL_SIZE_MODIFIER | N_VLP_per_light]
 *
 * Reads spheres.txt / squares.txt (nodof: planes.txt, falling back to squares.txt) /
 * triangles.txt / lights.txt from the current directory like the reference hosts do
 * (CLSuperPathTracer.c:261-264), renders with the C restatement and writes result.ppm.
 * Env: PT_SEEDS=a,b,c,d (default 1,2,3,4), PT_SPP (default 64), PT_OUT (default result.ppm),
 *      PT_THREADS.  Prints one `ORACLE_STATS {json}` line with time and work counters.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "oracle.h"

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s <base|lmem|nodof|grid|bidir> [w] [h] [cell_size_modifier | n_vlp]\n", argv[0]);
        return 2;
    }
    oracle_job J;
    memset(&J, 0, sizeof(J));
    const char *names[5] = {"base", "lmem", "nodof", "grid", "bidir"};
    J.variant = -1;
    for (int i = 0; i < 5; ++i)
        if (!strcmp(argv[1], names[i])) J.variant = i;
    if (J.variant < 0) { fprintf(stderr, "unknown variant %s\n", argv[1]); return 2; }
    J.width = argc > 2 ? atoi(argv[2]) : 512;
    J.height = argc > 3 ? atoi(argv[3]) : 512;
    float modifier = argc > 4 ? (float)atof(argv[4]) : 3.0f;
    J.spp = getenv("PT_SPP") ? atoi(getenv("PT_SPP")) : 64;
    J.nthreads = getenv("PT_THREADS") ? atoi(getenv("PT_THREADS")) : 0;
    unsigned long s[4] = {1, 2, 3, 4};
    if (getenv("PT_SEEDS")) sscanf(getenv("PT_SEEDS"), "%lu,%lu,%lu,%lu", &s[0], &s[1], &s[2], &s[3]);
    for (int i = 0; i < 4; ++i) J.seeds[i] = (uint32_t)s[i];

    int max_tris = J.variant == ORACLE_GRID ? 65536 : 512;   /* MAX_TRIANGLES of each host */
    if (getenv("PT_MAX_TRIANGLES")) max_tris = atoi(getenv("PT_MAX_TRIANGLES"));
    float *tris = (float *)malloc(sizeof(float) * 12 * (size_t)max_tris);
    if (oracle_parse_array("spheres.txt", J.spheres) < 0) { fprintf(stderr, "cannot open spheres.txt\n"); return 1; }
    const char *sq = "squares.txt";
    if (J.variant == ORACLE_NODOF) {
        FILE *f = fopen("planes.txt", "r");
        if (f) { fclose(f); sq = "planes.txt"; }
    }
    if (oracle_parse_array(sq, J.squares) < 0) { fprintf(stderr, "cannot open %s\n", sq); return 1; }
    J.ntriangles = oracle_parse_triangles("triangles.txt", tris, max_tris, J.box_min, J.box_max);
    if (J.ntriangles < 0) { fprintf(stderr, "cannot open triangles.txt\n"); return 1; }
    J.triangles = tris;
    J.nlights = oracle_parse_lights("lights.txt", J.lights);
    if (J.nlights < 0) { fprintf(stderr, "cannot open lights.txt\n"); return 1; }
    float fwd[4];
    oracle_camera(fwd, J.cam_up, J.cam_right, J.eye_offset);

    uint32_t *cell_start = NULL, *cell_refs = NULL;
    if (J.variant == ORACLE_GRID) {
        oracle_grid_dims(J.box_min, J.box_max, J.ntriangles, modifier, J.grid_res, J.cell_size);
        size_t ncells = (size_t)J.grid_res[0] * J.grid_res[1] * J.grid_res[2];
        cell_start = (uint32_t *)malloc(sizeof(uint32_t) * (ncells + 1));
        uint64_t total = oracle_build_grid(tris, J.ntriangles, J.box_min, J.grid_res, J.cell_size, 62, cell_start, NULL);
        cell_refs = (uint32_t *)malloc(sizeof(uint32_t) * (total ? total : 1));
        oracle_build_grid(tris, J.ntriangles, J.box_min, J.grid_res, J.cell_size, 62, cell_start, cell_refs);
        J.cell_start = cell_start;
        J.cell_refs = cell_refs;
        printf("Triangles grid size: %d x %d x %d\n", J.grid_res[0], J.grid_res[1], J.grid_res[2]);
    }
    printf("Number of triangles: %d\nNumber of lights: %d\n", J.ntriangles, J.nlights);

    uint8_t *img = (uint8_t *)malloc((size_t)J.width * J.height * 4);
    oracle_counters c;
    double t0 = now_ms();
    float *vpls = NULL;
    if (J.variant == ORACLE_BIDIR) {     /* light pass first, same seeds (CLSuperBidirectionalPathTracer.c:370-375); timed */
        int n_vlp = argc > 4 ? atoi(argv[4]) : 512;
        vpls = (float *)malloc(sizeof(float) * 4 * (size_t)(n_vlp > 0 ? n_vlp : 1) * (J.nlights > 0 ? J.nlights : 1));
        if (oracle_light_tracer(&J, n_vlp, vpls, NULL, NULL)) { fprintf(stderr, "oracle_light_tracer: bad job\n"); return 1; }
        J.vpls = vpls;
        J.nvpl = n_vlp * J.nlights;
    }
    if (oracle_render(&J, img, NULL, NULL, &c)) { fprintf(stderr, "oracle_render: bad job\n"); return 1; }
    double ms = now_ms() - t0;
    const char *out = getenv("PT_OUT") ? getenv("PT_OUT") : "result.ppm";
    if (oracle_save_pam(out, J.width, J.height, img)) { fprintf(stderr, "error writing %s\n", out); return 1; }
    printf("ORACLE_STATS {\"variant\": \"%s\", \"width\": %d, \"height\": %d, \"spp\": %d, \"ms\": %.3f, "
           "\"samples\": %llu, \"rays\": %llu, \"shadow_rays\": %llu, \"tri_tests\": %llu, \"cells_visited\": %llu, "
           "\"prim_tests\": %llu, \"contract\": %d}\n",
           names[J.variant], J.width, J.height, J.spp, ms, (unsigned long long)c.samples, (unsigned long long)c.rays,
           (unsigned long long)c.shadow_rays, (unsigned long long)c.tri_tests, (unsigned long long)c.cells_visited,
           (unsigned long long)c.prim_tests, oracle_contract_mode());
    free(img); free(tris); free(cell_start); free(cell_refs); free(vpls);
    return 0;
}
