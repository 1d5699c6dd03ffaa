/*
 * pthost_cli_metro.c — drop-in for CLSuperMetropolisPathTracer_vlpgrid/CLSuperMetropolisPathTracer.c:429-720: same argv
 * ([img_width] [img_height] [N_seedpaths_per_light] [mutation_rounds] [CELL_SIZE_MODIFIER]), scene files from the CWD, the
 * reference's stdout lines in its order, result.ppm.
 *
 * The reference host cannot run as written: it hands clCreateBuffer host pointers without a HOST_PTR flag, passes the VPL buffer
 * where lightTracer expects the seed paths (:579), and its kernel reads an uninitialised hit bound.  This
 * program runs the pipeline the host MEANS — seed paths -> Metropolis pass -> VPL bounding box -> VPL grid -> path tracer —
 * through libptcuda's FIX-mode kernels (include/ptcuda.h: pt_launch_metropolis_lighttracer).  Single GPU.
 * Environment: PT_SEEDS, PT_SPP, PT_ARITH, PT_DEVICE / OCL_DEVICE, PT_EXTRA_OUTPUT (as the other drop-ins).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "pthost.h"

int pth_cli_metropolis_main(int argc, char **argv) {
    int img_width = 512, img_height = 512, nseedpaths = 512, mutation_rounds = 8;
    float cell_size_modifier = 3.0f;
    printf("Usage: %s [img_width] [img_height] [N_seedpaths_per_light] [mutation_rounds] [CELL_SIZE_MODIFIER]\n"
           "Loads data from triangles.txt, lights.txt, spheres.txt and squares.txt\n", argv[0]);
    if (argc > 1) img_width = atoi(argv[1]);
    if (argc > 2) img_height = atoi(argv[2]);
    if (argc > 3) nseedpaths = atoi(argv[3]);
    if (argc > 4) mutation_rounds = atoi(argv[4]);
    if (argc > 5) cell_size_modifier = (float)atof(argv[5]);

    /* select_platform / select_device / create_* of ocl_boiler.h */
    printf("number of platforms: %u\n", 1u);
    printf("selected platform %d: %s\n", 0, "NVIDIA CUDA (libptcuda, sm_100a)");
    const int dev = pt_select_device();
    pt_ctx ctx = pt_create(dev);
    if (!ctx) pt_check(1, "create the CUDA context");
    time_t now = time(NULL);
    printf("compiling:\n// %s#include \"%s\"\n", ctime(&now), "metropolispathtracer.ocl");
    printf("=== BUILD LOG ===\n%s\n=========\n", "kernels are precompiled CUDA for sm_100a (libptcuda.so); nothing to build\n");

    uint32_t seeds[4];
    pth_seeds(seeds);
    printf("Seeds: %d, %d, %d, %d\n", (int)seeds[0], (int)seeds[1], (int)seeds[2], (int)seeds[3]);
    long data_size = (long)img_width * img_height * 4;
    printf("Processing image %dx%d with data size %ld bytes\n", img_width, img_height, data_size);
    pt_camera cam;
    pth_camera(&cam);
    printf("Cam values:\nCam_forward %f %f %f\nCam_up %f %f %f\nCam_right %f %f %f\n eye_offset %f %f %f\n",
           cam.cam_forward[0], cam.cam_forward[1], cam.cam_forward[2], cam.cam_up[0], cam.cam_up[1], cam.cam_up[2],
           cam.cam_right[0], cam.cam_right[1], cam.cam_right[2], cam.eye_offset[0], cam.eye_offset[1], cam.eye_offset[2]);

    pt_scene scene;
    memset(&scene, 0, sizeof(scene));
    if (pth_parse_bitmap("spheres.txt", scene.spheres) < 0) pt_check(1, "open spheres.txt");
    if (pth_parse_bitmap("squares.txt", scene.squares) < 0) pt_check(1, "open squares.txt");
    int max_triangles = 512;
    if (getenv("PT_MAX_TRIANGLES")) max_triangles = atoi(getenv("PT_MAX_TRIANGLES"));
    float *tris = NULL;
    float box_min[4], box_max[4];
    scene.ntriangles = pth_parse_triangles("triangles.txt", max_triangles, &tris, box_min, box_max);
    if (scene.ntriangles < 0) pt_check(1, "open triangles.txt");
    scene.triangles = tris;
    scene.nlights = pth_parse_lights("lights.txt", scene.lights, 0);
    if (scene.nlights < 0) pt_check(1, "open lights.txt");
    const int n_vlp = nseedpaths * scene.nlights * 4;                  /* :536 */
    printf("Number of triangles: %d\n", scene.ntriangles);
    printf("Number of lights: %d\n", scene.nlights);
    printf("Mutation rounds: %d\n", mutation_rounds);
    pt_set_scene(ctx, &scene);

    const char *ar = getenv("PT_ARITH");
    const int arith = (ar && !strcmp(ar, "separate")) ? PT_ARITH_SEPARATE : PT_ARITH_FMA;
    /* :579-581 lightTracer + MetropolisLightTracer (FIX mode, one launch) */
    pt_event light_evt = pt_launch_metropolis_lighttracer(ctx, nseedpaths, seeds, mutation_rounds, arith);
    if (!light_evt) pt_check(1, "metropolis light tracer");
    /* :583-604 reduction(): the reference prints its launch geometry (lws 256; _nwg pass when more than one group) */
    {
        const int lws = 256, nwg = (n_vlp + lws - 1) / lws;
        printf("gws: %d, lws: %d\n", nwg * lws, lws);
        if (nwg > 1) printf("gws: %d, lws: %d\n", lws, lws);
    }
    float vmin[4], vmax[4];
    if (pt_vlp_bounds(ctx, vmin, vmax)) pt_check(1, "VLP bounding box");
    printf("VLPs bounding box values:\nvmax: %f %f %f, vmin: %f %f %f\n", vmax[0], vmax[1], vmax[2], vmin[0], vmin[1], vmin[2]);
    pt_grid vg;
    memset(&vg, 0, sizeof(vg));
    pth_grid_dims(vmin, vmax, n_vlp, cell_size_modifier, &vg);         /* :628-636: the triangle-grid formula on N_VLP */
    printf("VLPs grid size: %d x %d x %d\n", vg.res[0], vg.res[1], vg.res[2]);
    pt_event grid_evt = pt_build_vlp_grid(ctx, &vg);
    if (!grid_evt) pt_check(1, "init VLPs grid");

    pt_render_params rp;
    memset(&rp, 0, sizeof(rp));
    rp.variant = PT_VARIANT_VLPGRID;
    rp.width = img_width; rp.height = img_height;
    rp.spp = getenv("PT_SPP") ? atoi(getenv("PT_SPP")) : 64;
    memcpy(rp.seeds, seeds, sizeof(seeds));
    rp.kernel = PT_KERNEL_AUTO; rp.scene_mem = PT_SCENE_AUTO; rp.arith = arith;
    pt_event render_evt = pt_launch_pathtracer(ctx, &cam, &rp);
    if (!render_evt) pt_check(1, "path tracer");
    pt_event read_evt = NULL;
    void *pixels = pt_map_render(ctx, &read_evt);

    const char *image_name = "result.ppm";
    if (pth_save_pam(image_name, img_width, img_height, pixels) != 0) {
        fprintf(stderr, "error writing %s\n", image_name);
        exit(1);
    } else
        printf("\nSuccessfully created render image %s in the current directory\n\n", image_name);
    const char *extra = getenv("PT_EXTRA_OUTPUT");
    if (extra && strstr(extra, "png") && pth_save_png("result.png", img_width, img_height, pixels)) pt_check(1, "write result.png");
    if (extra && strstr(extra, "ppm") && pth_save_ppm("result_p6.ppm", img_width, img_height, pixels)) pt_check(1, "write result_p6.ppm");

    /* both light kernels run in ONE launch here: its time is reported on the first line, the second one gets 0 of it */
    const double light_ms = pt_runtime_ms(light_evt), grid_ms = pt_runtime_ms(grid_evt), render_ms = pt_runtime_ms(render_evt),
                 read_ms = pt_runtime_ms(read_evt);
    const int ncells = vg.res[0] * vg.res[1] * vg.res[2];
    printf("light paths random sampling : %d random light paths in %gms: %g GB/s\n", nseedpaths * scene.nlights, light_ms,
           nseedpaths * scene.nlights * 16.0 * 4 / 1.0e6 / light_ms);
    printf("light paths metropolis sampling : %d virtual lights in %gms: %g GB/s\n", n_vlp, 0.0, 0.0);
    printf("VLPs min/max reduction (compute bounding box) : %d virtual lights in %gms: %g GB/s\n", n_vlp, 0.0, 0.0);
    printf("Read VLPs bounding box in %gms: %g GB/s\n", 0.0, 0.0);
    printf("init VLPs grid : %d cells in %gms: %g GB/s\n", ncells, grid_ms, 128.0 * ncells / 1.0e6 / grid_ms);
    printf("rendering : %d pixels in %gms: %g GB/s\n", img_width * img_height, render_ms, data_size / 1.0e6 / render_ms);
    printf("read render data : %ld uchar in %gms: %g GB/s\n", data_size, read_ms, data_size / 1.0e6 / read_ms);
    printf("\nTotal time: %g ms.\n", light_ms + grid_ms + render_ms + read_ms);

    pt_release_event(light_evt); pt_release_event(grid_evt); pt_release_event(render_evt); pt_release_event(read_evt);
    free(tris);
    pt_destroy(ctx);
    return 0;
}
