/*
 * ref_kernels.cpp — compiles ONE reference pathtracer.ocl (selected by
 * -DREF_VARIANT=0..4 and -DREF_OCL_GEN="<generated file>") as host C++ and
 * registers its kernels with refrt.  TEST INFRASTRUCTURE ONLY.
 *
 *   REF_VARIANT 0  CLSuperPathTracer/pathtracer.ocl              (base)
 *   REF_VARIANT 1  CLSuperPathTracer_lmem/pathtracer.ocl         (lmem)
 *   REF_VARIANT 2  CLSuperPathTracer_lmem_NoDoF/pathtracer.ocl   (nodof)
 *   REF_VARIANT 3  CLSuperPathTracer_trianglegrid/pathtracer.ocl (grid)
 *   REF_VARIANT 4  CLSuperBidirectionalPathTracer/bidirectionalpathtracer.ocl (bidir)
 *   REF_VARIANT 5  CLSuperMetropolisPathTracer_vlpgrid/metropolispathtracer.ocl (VLP bounding box, VLP grid and pathTracer kernels)
 *
 * REF_OCL_GEN is the reference source with vector literals rewritten by ocl2cpp.py; the Makefile
 * pipes it in ("/dev/stdin"), so no copy of it is written anywhere.  Argument
 * indices below follow the clSetKernelArg sequences of the reference hosts:
 *   base  CLSuperPathTracer/CLSuperPathTracer.c:153-177
 *   lmem  CLSuperPathTracer_lmem/CLSuperPathTracer.c:154-186
 *   nodof CLSuperPathTracer_lmem_NoDoF/CLSuperPathTracer.c:155-187, 207-210
 *   grid  CLSuperPathTracer_trianglegrid/CLSuperPathTracer.c:288-297, 337-374
 *   bidir CLSuperBidirectionalPathTracer/CLSuperBidirectionalPathTracer.c:154-177 (lightTracer), 200-235 (pathTracer)
 *
 * It also exports a few `ref_probe_*` C functions that call the reference's
 * inline device functions directly (RNG, TraceRay) so that unit-level golden
 * vectors can be produced from the real reference code.
 */
#include "clshim.h"

namespace ocl {
#include REF_OCL_GEN
}  // namespace ocl

#undef kernel
#undef global
#undef local
#undef constant
#undef restrict

#include "refrt.h"

using namespace ocl;

static inline uint4 seeds_arg(const RefLaunch &L, int i) {
    /* decoded once per work-item: cheap next to 64 samples of path tracing */
    uint4 s = L.val<uint4>(i);
    static thread_local bool have = false;
    static thread_local uint32_t ov[4];
    static thread_local uint32_t src[4];
    if (!have || src[0] != s.x || src[1] != s.y || src[2] != s.z || src[3] != s.w) {
        src[0] = s.x; src[1] = s.y; src[2] = s.z; src[3] = s.w;
        ov[0] = s.x; ov[1] = s.y; ov[2] = s.z; ov[3] = s.w;
        refrt_override_seeds(ov);
        have = true;
    }
    return uint4(ov[0], ov[1], ov[2], ov[3]);
}

#if REF_VARIANT == 0
static void tramp_pathTracer(const RefLaunch &L) {
    ocl::pathTracer((uchar4 *)L.mem(0), (int *)L.mem(1), (int *)L.mem(2), (const Triangle *)L.mem(3), L.val<int>(4),
                    (float4 *)L.mem(5), L.val<int>(6), L.val<float4>(7), L.val<float4>(8), L.val<float4>(9),
                    L.val<float4>(10), seeds_arg(L, 11));
}
const RefKernelDesc ref_kernel_table[] = {{"pathTracer", 12, tramp_pathTracer}, {nullptr, 0, nullptr}};
#elif REF_VARIANT == 1
static void tramp_pathTracer(const RefLaunch &L) {
    ocl::pathTracer((uchar4 *)L.mem(0), (const int *)L.mem(1), (const int *)L.mem(2), (const Triangle *)L.mem(3),
                    L.val<int>(4), (const float4 *)L.mem(5), L.val<int>(6), L.val<float4>(7), L.val<float4>(8),
                    L.val<float4>(9), L.val<float4>(10), seeds_arg(L, 11), (int *)L.local(12), (int *)L.local(13),
                    (Triangle *)L.local(14), (float4 *)L.local(15));
}
const RefKernelDesc ref_kernel_table[] = {{"pathTracer", 16, tramp_pathTracer}, {nullptr, 0, nullptr}};
#elif REF_VARIANT == 2
static void tramp_pathTracer(const RefLaunch &L) {
    ocl::pathTracer((float4 *)L.mem(0), (int *)L.mem(1), (int *)L.mem(2), (Triangle *)L.mem(3), L.val<int>(4),
                    (float4 *)L.mem(5), L.val<int>(6), L.val<float4>(7), L.val<float4>(8), L.val<float4>(9),
                    L.val<float4>(10), seeds_arg(L, 11), (int *)L.local(12), (int *)L.local(13),
                    (Triangle *)L.local(14), (float4 *)L.local(15));
}
static void tramp_reduce(const RefLaunch &L) {
    ocl::reduce4img_lmem((const float4 *)L.mem(0), (uchar4 *)L.mem(1), (float4 *)L.local(2));
}
const RefKernelDesc ref_kernel_table[] = {
    {"pathTracer", 16, tramp_pathTracer}, {"reduce4img_lmem", 3, tramp_reduce}, {nullptr, 0, nullptr}};
#elif REF_VARIANT == 3
static void tramp_pathTracer(const RefLaunch &L) {
    ocl::pathTracer((uchar4 *)L.mem(0), (const int *)L.mem(1), (const int *)L.mem(2), (const Triangle *)L.mem(3),
                    L.val<int>(4), L.val<Box>(5), (const Cell *)L.mem(6), L.val<int4>(7), L.val<float4>(8),
                    (const float4 *)L.mem(9), L.val<int>(10), L.val<float4>(11), L.val<float4>(12), L.val<float4>(13),
                    L.val<float4>(14), seeds_arg(L, 15), (int *)L.local(16), (int *)L.local(17), (float4 *)L.local(18));
}
static void tramp_initgrid(const RefLaunch &L) {
    ocl::initTrianglesGrid((Cell *)L.mem(0), (const Triangle *)L.mem(1), L.val<float4>(2), L.val<int4>(3),
                           L.val<float4>(4));
}
static void tramp_printgrid(const RefLaunch &L) { ocl::printTrianglesGrid((const Cell *)L.mem(0)); }
const RefKernelDesc ref_kernel_table[] = {{"pathTracer", 19, tramp_pathTracer},
                                          {"initTrianglesGrid", 5, tramp_initgrid},
                                          {"printTrianglesGrid", 1, tramp_printgrid},
                                          {nullptr, 0, nullptr}};
#elif REF_VARIANT == 4
static void tramp_pathTracer(const RefLaunch &L) {
    ocl::pathTracer((uchar4 *)L.mem(0), (const int *)L.mem(1), (const int *)L.mem(2), (const Triangle *)L.mem(3),
                    L.val<int>(4), (const float4 *)L.mem(5), L.val<int>(6), (const float4 *)L.mem(7), L.val<int>(8),
                    L.val<float4>(9), L.val<float4>(10), L.val<float4>(11), L.val<float4>(12), seeds_arg(L, 13),
                    (int *)L.local(14), (int *)L.local(15), (Triangle *)L.local(16), (float4 *)L.local(17));
}
static void tramp_lightTracer(const RefLaunch &L) {
    ocl::lightTracer((const int *)L.mem(0), (const int *)L.mem(1), (const Triangle *)L.mem(2), L.val<int>(3),
                     (const float4 *)L.mem(4), L.val<int>(5), (float4 *)L.mem(6), seeds_arg(L, 7), (int *)L.local(8),
                     (int *)L.local(9), (Triangle *)L.local(10), (float4 *)L.local(11));
}
const RefKernelDesc ref_kernel_table[] = {
    {"pathTracer", 18, tramp_pathTracer}, {"lightTracer", 12, tramp_lightTracer}, {nullptr, 0, nullptr}};
#elif REF_VARIANT == 5
/* CLSuperMetropolisPathTracer_vlpgrid/metropolispathtracer.ocl: only the kernels that are pure functions of a VLP buffer
 * are registered (argument order: CLSuperMetropolisPathTracer.c:262-296 reduction(), :298-321 initVLPsGrid(), :186-227
 * pathTracer() — a function of scene + VLP buffer + VLP grid); the Metropolis light tracer of that program is compiled
 * but not exposed. */
static void tramp_pathTracer(const RefLaunch &L) {
    ocl::pathTracer((uchar4 *)L.mem(0), (const int *)L.mem(1), (const int *)L.mem(2), (const Triangle *)L.mem(3), L.val<int>(4),
                    (const float4 *)L.mem(5), L.val<int>(6), (const Cell *)L.mem(7), L.val<float4>(8), L.val<float4>(9),
                    L.val<int4>(10), (const float4 *)L.mem(11), L.val<int>(12), L.val<float4>(13), L.val<float4>(14),
                    L.val<float4>(15), L.val<float4>(16), seeds_arg(L, 17), (int *)L.local(18), (int *)L.local(19),
                    (float4 *)L.local(20));
}
static void tramp_reduce_minmax(const RefLaunch &L) {
    ocl::reduceMinAndMax_lmem((float4 *)L.mem(0), (float8 *)L.mem(1), (float8 *)L.local(2), L.val<int>(3));
}
static void tramp_reduce_minmax_nwg(const RefLaunch &L) {
    ocl::reduceMinAndMax_lmem_nwg((float8 *)L.mem(0), (float8 *)L.mem(1), (float8 *)L.local(2), L.val<int>(3));
}
static void tramp_init_vlps_grid(const RefLaunch &L) {
    ocl::initVLPsGrid((Cell *)L.mem(0), (const float4 *)L.mem(1), L.val<float4>(2), L.val<int4>(3), L.val<float4>(4));
}
#if defined(REF_METRO_FIX)
/* FIX build only (oracle/Makefile: `float t = 1e9;` in VerifyIntersection): the seed-path tracer and the Metropolis light tracer,
 * argument order of CLSuperMetropolisPathTracer_vlpgrid/CLSuperMetropolisPathTracer.c:186-260 */
static void tramp_lightTracer(const RefLaunch &L) {
    ocl::lightTracer((const int *)L.mem(0), (const int *)L.mem(1), (const Triangle *)L.mem(2), L.val<int>(3), (const float4 *)L.mem(4),
                     L.val<int>(5), (Path *)L.mem(6), seeds_arg(L, 7), (int *)L.local(8), (int *)L.local(9), (float4 *)L.local(10));
}
static void tramp_metropolis(const RefLaunch &L) {
    ocl::MetropolisLightTracer((const int *)L.mem(0), (const int *)L.mem(1), (const Triangle *)L.mem(2), L.val<int>(3),
                               (const float4 *)L.mem(4), L.val<int>(5), (Path *)L.mem(6), (float16 *)L.mem(7), seeds_arg(L, 8),
                               L.val<int>(9), (int *)L.local(10), (int *)L.local(11), (float4 *)L.local(12));
}
#endif
const RefKernelDesc ref_kernel_table[] = {{"reduceMinAndMax_lmem", 4, tramp_reduce_minmax},
                                          {"reduceMinAndMax_lmem_nwg", 4, tramp_reduce_minmax_nwg},
                                          {"initVLPsGrid", 5, tramp_init_vlps_grid},
                                          {"pathTracer", 21, tramp_pathTracer},
#if defined(REF_METRO_FIX)
                                          {"lightTracer", 11, tramp_lightTracer},
                                          {"MetropolisLightTracer", 13, tramp_metropolis},
#endif
                                          {nullptr, 0, nullptr}};
#else
#error "REF_VARIANT must be 0..5"
#endif

/* ------------------------------------------------------------------ probes */
extern "C" {

int ref_variant(void) { return REF_VARIANT; }

/* Runs MWC64XVEC2_Seeding (pathtracer.ocl:37-41) for work-item `gid` of a 1-D
 * range, then `nsteps` calls of MWC64XVEC2(&s, 0, 1) (pathtracer.ocl:12-23).
 * out_f[2*k..] = returned floats, out_state = {x.x, x.y, c.x, c.y} afterwards. */
void ref_probe_rng(const uint32_t seeds[4], uint32_t gid, int nsteps, float *out_f, uint32_t out_state[4]) {
    size_t gsz[3] = {(size_t)gid + 1, 1, 1}, one[3] = {1, 1, 1};
    WorkItem wi;
    wi.gid[0] = gid; wi.gid[1] = 0; wi.gid[2] = 0;
    wi.lid[0] = wi.lid[1] = wi.lid[2] = 0;
    wi.grp[0] = gid; wi.grp[1] = wi.grp[2] = 0;
    wi.gsz = gsz; wi.lsz = one; wi.ngrp = gsz;
    WorkItem *saved = refrt_wi;
    refrt_wi = &wi;
    mwc64xvec2_state_t s;
    MWC64XVEC2_Seeding(&s, uint4(seeds[0], seeds[1], seeds[2], seeds[3]));
    for (int k = 0; k < nsteps; ++k) {
        float2 r = MWC64XVEC2(&s, 0.0f, 1.0f);
        out_f[2 * k] = r.x;
        out_f[2 * k + 1] = r.y;
    }
    out_state[0] = s.x.x; out_state[1] = s.x.y; out_state[2] = s.c.x; out_state[3] = s.c.y;
    refrt_wi = saved;
}

uint32_t ref_probe_randomize_id(uint32_t id) { return randomizeId(id); }

#if REF_VARIANT != 3
/* Calls the reference TraceRay (pathtracer.ocl:48-137 / lmem :48-135) on one ray.
 * t_inout: base ignores the input value (it resets to 1e9 itself). */
int ref_probe_trace_ray(const float o[3], const float d[3], float *t_inout, float n_out[3], const int32_t spheres[9],
                        const int32_t squares[9], const float *tris12, int ntris) {
    float4 origin(o[0], o[1], o[2], 0.0f), dir(d[0], d[1], d[2], 0.0f), normal(0.0f, 0.0f, 0.0f, 0.0f);
    int sp[9], sq[9];
    for (int i = 0; i < 9; ++i) { sp[i] = spheres[i]; sq[i] = squares[i]; }
    std::vector<Triangle> tris(ntris > 0 ? ntris : 1); /* 16-byte aligned copy */
    if (ntris > 0) std::memcpy((void *)tris.data(), tris12, sizeof(Triangle) * (size_t)ntris);
    int m = TraceRay(origin, dir, t_inout, &normal, sp, sq, tris.data(), ntris);
    n_out[0] = normal.x; n_out[1] = normal.y; n_out[2] = normal.z;
    return m;
}
#endif

#if defined(REF_METRO_FIX)
/* The reference's Mutate (metropolispathtracer.ocl:238-294; FIX build) applied `rounds` times to one path, seeded as work-item
 * `gid` of the Metropolis kernels: the kernel itself keeps the mutated path in a private copy, so this is the only way to see it.
 * path: the reference's Path layout, 20 words {float4 v[4]; uint length; pad}. */
void ref_probe_mutate(const uint32_t seeds[4], uint32_t gid, const int32_t spheres[9], const int32_t squares[9], const float *tris12,
                      int ntris, const float origin[3], uint32_t path[20], int rounds) {
    size_t gsz[3] = {(size_t)gid + 1, 1, 1}, one[3] = {1, 1, 1};
    WorkItem wi;
    wi.gid[0] = gid; wi.gid[1] = 0; wi.gid[2] = 0;
    wi.lid[0] = wi.lid[1] = wi.lid[2] = 0;
    wi.grp[0] = gid; wi.grp[1] = wi.grp[2] = 0;
    wi.gsz = gsz; wi.lsz = one; wi.ngrp = gsz;
    WorkItem *saved = refrt_wi;
    refrt_wi = &wi;
    mwc64xvec2_state_t rng;
    MWC64XVEC2_Seeding(&rng, uint4(seeds[0], seeds[1], seeds[2], seeds[3]));
    int sp[9], sq[9];
    for (int i = 0; i < 9; ++i) { sp[i] = spheres[i]; sq[i] = squares[i]; }
    std::vector<Triangle> tris(ntris > 0 ? ntris : 1);
    if (ntris > 0) std::memcpy((void *)tris.data(), tris12, sizeof(Triangle) * (size_t)ntris);
    Path p;
    std::memcpy((void *)&p, path, 80);
    const float4 o(origin[0], origin[1], origin[2], 0.0f);
    for (int m = 0; m < rounds; ++m) Mutate(&p, o, sp, sq, tris.data(), ntris, rng);
    std::memcpy(path, (const void *)&p, 80);
    refrt_wi = saved;
}
#endif

}  // extern "C"
