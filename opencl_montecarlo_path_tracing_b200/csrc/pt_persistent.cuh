// pt_persistent.cuh — per-lane ray STATE MACHINE kernels (PT_KERNEL_PERSISTENT).
//
// The sequential megakernel (pt_mega.cuh) makes every lane of a warp wait while its neighbours trace
// their shadow rays, and inlines TraceRay twice (primary + shadow).  Here each lane carries a small
// state {primary pending | shadow ray for light l pending} and every loop iteration runs ONE shared
// TraceRay body for whatever ray each lane needs next:
//   * a lane whose sample ended in the sky (1 ray) immediately starts its next sample while its
//     neighbours are still on shadow rays (up to 3 rays/sample) -> no idle lanes inside a pixel;
//   * lanes that finished their pixel fetch a new one from a global work counter with one
//     warp-aggregated atomicAdd (__ballot_sync / __popc) -> ray REGENERATION, persistent CTAs;
//   * one TraceRay body instead of two -> half the instruction-cache footprint.
// Bit-exactness is untouched: each pixel still consumes its own RNG stream in order;
// only the interleaving of independent pixels on a lane changes.
#pragma once
#include <numeric>

#include "pt_host.h"

namespace pt {

struct Lane {
    Rng rng;
    int px, py;        // pixel coordinates used by the camera (NoDoF: the pixel, not the sample id)
    int phase;         // 0: primary ray to be generated, 1: shadow ray for light `l` in flight
    int l;             // light index of the shadow ray in flight / next light to visit
    int mat;           // material of the primary hit (1 floor, 3 diffuse, 4 facing ratio)
    V3 o, d;           // ray to trace next
    float t;           // running *t of the reference (lmem family: carried across the sample's rays)
    V3 X, n;           // primary hit point and normal
    float illum, lam;  // accumulated illumination, Lambert factor of the light in flight
    float matf;        // m == 1: checker parity (0/1); m == 4: facing ratio
};

// Visit lights from L.l on until one needs a shadow ray (returns true, ray set up) or none is left.
template <bool FMA, bool CARRY>
PT_DEV bool next_shadow_ray(const AnalyticParams &AP, Lane &L, Counters &cnt) {
    while (L.l < AP.nlights) {
        float r0, r1;
        rng_next(L.rng, r0, r1);                       // drawn before any skip (base:168)
        const float4 Lt = AP.lights[L.l];
        if (!CARRY && Lt.w == 0.0f) { L.l++; continue; }   // base:171 only
        V3 ld; float lam;
        light_dir<FMA>(Lt, r0, r1, L.X, L.n, ld, lam);
        if (lam < 0.0f) { L.l++; continue; }
        L.o = L.X; L.d = ld; L.lam = lam; L.phase = 1;
        cnt.shadow++;
        return true;
    }
    return false;
}

template <bool FMA>
PT_DEV V3 finish_material(const Lane &L) {
    typedef Ar<FMA> A;
    float illum = L.illum;
    if (illum > 1.0f) illum = 1.0f;
    illum = A::mul(illum, 0.25f);
    if (L.mat == 1) {
        float i3 = A::mul(3.0f, illum);
        return L.matf != 0.0f ? mk3(i3, illum, illum) : mk3(i3, i3, i3);
    }
    if (L.mat == 3) { float i2 = A::mul(2.0f, illum); return mk3(i2, A::mul(3.0f, illum), i2); }
    return mk3(L.matf, L.matf, L.matf);
}

// One state-machine step for an active lane: generate (if needed), trace, react.
// Returns true when a Sample() completed; its colour is in `out`.
template <bool FMA, bool CARRY, bool GRID>
PT_DEV bool lane_step(const LaunchArgs &P, const SceneBlock *S, Lane &L, bool active, V3 &out, Counters &cnt) {
    typedef Ar<FMA> A;
    if (active && L.phase == 0) {
        camera_ray<FMA>(P.cam, L.rng, L.px, L.py, L.o, L.d);
        L.t = 1e9f;                                    // lmem:155 (base resets inside TraceRay anyway)
    }
    int hit = HIT_NONE;
    if (active) hit = trace_ray<FMA, CARRY, GRID>(P.ap, S, P.grid, L.o, L.d, L.t, cnt);
    if (!active) return false;
    if (L.phase == 0) {
        cnt.samples++;
        if (hit == HIT_NONE) { out = shade_sky<FMA>(L.d); return true; }
        L.mat = hit_material(hit);
        L.n = hit_normal<FMA, GRID>(P.ap, S, P.grid, hit, L.o, L.d, L.t);
        L.X = A::vmadd(L.d, L.t, L.o);
        L.illum = 0.0f;
        L.matf = 0.0f;
        if (L.mat == 1) {                              // checkerboard parity (base:196-197)
            float yx = A::mul(L.X.x, 0.2f), yy = A::mul(L.X.y, 0.2f);
            L.matf = (f2i_rz_sat(A::add(ceilf(yx), ceilf(yy))) & 1) ? 1.0f : 0.0f;
        } else if (L.mat == 4) {                       // facing ratio (base:203-205)
            float fr = A::dot(L.n, mk3(-L.d.x, -L.d.y, -L.d.z));
            L.matf = 0.0f < fr ? fr : 0.0f;
        }
        L.l = 0;
    } else {
        if (hit == HIT_NONE) L.illum = light_add<FMA>(P.ap.lights[L.l], L.X, L.lam, L.illum);
        L.l++;
    }
    if (next_shadow_ray<FMA, CARRY>(P.ap, L, cnt)) return false;
    L.phase = 0;
    out = finish_material<FMA>(L);
    return true;
}

// work item -> pixel: items are numbered tile-major over 8x4 pixel tiles so that the 32 items a warp
// holds at any time stay spatially close (coherent rays) even after regeneration.
PT_DEV bool item_to_pixel(const LaunchArgs &P, uint32_t w, int &i, int &j, uint32_t nitems = 0) {
    // Scenes with brute-force triangles: the few pixels that see the mesh cost ~100x more than the rest and
    // sit next to each other.  A multiplicative permutation of the item order deals them out one or two per
    // warp, so (a) the cooperative triangle scan applies (few lanes need it) and (b) no warp inherits a long
    // serial chain of heavy pixels.  Which lane renders which pixel never affects results.
    if (P.scatter_mul && nitems) w = (uint32_t)(((unsigned long long)w * P.scatter_mul) % nitems);
    const uint32_t tiles_x = (uint32_t)(P.W + 7) >> 3;
    const uint32_t tile = w >> 5, lit = w & 31;
    const uint32_t ty = tile / tiles_x, tx = tile - ty * tiles_x;
    i = (int)(tx * 8 + (lit & 7));
    const int vr = (int)(ty * 4 + (lit >> 3));
    j = map_row(P, vr);
    return i < P.W && vr < P.nrows && j < P.row_end;
}

template <int VARIANT, bool FMA, int MEM, bool REGEN>
__global__ void __launch_bounds__(128, 6) k_sm_pixel(const __grid_constant__ LaunchArgs P, uint32_t nitems, uint32_t *work_counter) {
    constexpr bool CARRY = VARIANT != PT_VARIANT_BASE;
    constexpr bool GRID = VARIANT == PT_VARIANT_GRID;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SceneBlock *S = (MEM == PT_SCENE_SMEM) ? stage_scene_smem(P, smem_raw) : &c_scene;
    const unsigned lane = threadIdx.x & 31;
    Counters cnt = {0, 0, 0, 0, 0, 0};
    Lane L;
    L.phase = 0; L.l = 0; L.mat = 0; L.illum = 0.f; L.lam = 0.f; L.matf = 0.f; L.t = 1e9f;
    L.o = L.d = L.X = L.n = mk3(0.f, 0.f, 0.f);
    float cx = P.c0, cy = P.c0, cz = P.c0;
    int s = 0;
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    bool have = false;      // lane owns a valid pixel
    bool want = true;       // lane needs a (new) work item
    bool fetch = false;     // ... and must first draw a fresh index from the global counter
    for (;;) {
        // ---- (re)generation: hand out work items until every wanting lane has a valid pixel or none is left
        while (__any_sync(0xffffffffu, want)) {
            if (REGEN) {
                const unsigned need = __ballot_sync(0xffffffffu, want && fetch);
                if (need) {
                    const int leader = __ffs(need) - 1;
                    uint32_t base = 0;
                    if ((int)lane == leader) base = atomicAdd(work_counter, (uint32_t)__popc(need));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (want && fetch) w = base + __popc(need & ((1u << lane) - 1u));
                    fetch = false;
                }
            }
            if (want) {
                int i, j;
                if (w >= nitems) {
                    want = false;                      // queue exhausted
                } else if (item_to_pixel(P, w, i, j, nitems)) {
                    L.px = i; L.py = j;
                    L.rng = rng_seed(P.seeds, (uint32_t)(j * P.W + i));
                    L.phase = 0; s = 0; cx = cy = cz = P.c0;
                    have = true; want = false;
                } else if (REGEN) {
                    fetch = true;                      // item lies outside the image: draw another
                } else {
                    want = false;
                }
            }
        }
        if (!__any_sync(0xffffffffu, have)) break;
        V3 c;
        if (lane_step<FMA, CARRY, GRID>(P, S, L, have, c, cnt)) {
            cx = Ar<FMA>::madd(c.x, P.scale, cx);
            cy = Ar<FMA>::madd(c.y, P.scale, cy);
            cz = Ar<FMA>::madd(c.z, P.scale, cz);
            if (++s == P.spp) {
                const size_t pix = (size_t)L.py * P.W + L.px;
                P.rgba[pix] = pack_rgba8_rz(cx, cy, cz, P.alpha);
                if (P.accum) P.accum[pix] = make_float4(cx, cy, cz, P.alpha);
                if (P.rng_out) P.rng_out[pix] = make_uint4(L.rng.x0, L.rng.x1, L.rng.c0, L.rng.c1);
                have = false;
                want = REGEN;       // fetched at the top of the loop with one warp-aggregated atomicAdd
                fetch = REGEN;
            }
        }
    }
    if (GRID) cnt.btests = cnt.gtri;        // trace_grid tests every record of the cells it visits
    flush_counters(P, cnt, GRID ? 0 : S->ntri_counted, P.ap.nsq + P.ap.nsp);
}

// NoDoF: one warp per pixel (the 8x8 reduction tree lives in one warp), lanes run their two samples
// through the state machine; warps are persistent and stride over the pixels.
template <bool FMA, int MEM>
__global__ void __launch_bounds__(256, 4) k_sm_nodof(const __grid_constant__ LaunchArgs P, uint32_t npix_items) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SceneBlock *S = (MEM == PT_SCENE_SMEM) ? stage_scene_smem(P, smem_raw) : &c_scene;
    const int lane = threadIdx.x & 31;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    Counters cnt = {0, 0, 0, 0, 0, 0};
    const uint32_t tiles_x = (uint32_t)(P.W + 3) >> 2;    // pixels are walked in 4x2 tiles
    for (uint32_t item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < npix_items; item += warps_total) {
        const uint32_t tile = item >> 3, pit = item & 7;
        const uint32_t ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int px = (int)(tx * 4 + (pit & 3));
        const int vr = (int)(ty * 2 + (pit >> 2));
        const int py = map_row(P, vr);
        if (!(px < P.W && vr < P.nrows && py < P.row_end)) continue;    // warp-uniform
        Lane L;
        L.px = px; L.py = py; L.phase = 0; L.l = 0; L.mat = 0; L.illum = 0.f; L.lam = 0.f; L.matf = 0.f; L.t = 1e9f;
        L.o = L.d = L.X = L.n = mk3(0.f, 0.f, 0.f);
        float ax = 0.f, ay = 0.f, az = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
        int h = 0;              // which of the lane's two samples is running
        bool fresh = true;      // sample h has not been seeded yet
        while (__any_sync(0xffffffffu, h < 2)) {
            const bool active = h < 2;
            if (active && fresh) {
                const int li = lane + 32 * h;                       // local id inside the 8x8 group
                const int gi = 8 * px + (li & 7), gj = 8 * py + (li >> 3);
                L.rng = rng_seed(P.seeds, (uint32_t)(gj * (8 * P.W) + gi));
                L.phase = 0;
                fresh = false;
            }
            V3 c;
            if (lane_step<FMA, true, false>(P, S, L, active, c, cnt)) {
                if (h == 0) { ax = __fmul_rn(c.x, 3.5f); ay = __fmul_rn(c.y, 3.5f); az = __fmul_rn(c.z, 3.5f); }
                else        { bx = __fmul_rn(c.x, 3.5f); by = __fmul_rn(c.y, 3.5f); bz = __fmul_rn(c.z, 3.5f); }
                if (P.rng_out) {
                    const int li = lane + 32 * h;
                    const int gi = 8 * px + (li & 7), gj = 8 * py + (li >> 3);
                    P.rng_out[(size_t)gj * (8 * P.W) + gi] = make_uint4(L.rng.x0, L.rng.x1, L.rng.c0, L.rng.c1);
                }
                ++h;
                fresh = true;
            }
        }
        // nodof:253-274 reduction tree: li += li+32, then +16, +8, +4, +2, +1
        float sx = __fadd_rn(ax, bx), sy = __fadd_rn(ay, by), sz = __fadd_rn(az, bz);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sx = __fadd_rn(sx, __shfl_down_sync(0xffffffffu, sx, off));
            sy = __fadd_rn(sy, __shfl_down_sync(0xffffffffu, sy, off));
            sz = __fadd_rn(sz, __shfl_down_sync(0xffffffffu, sz, off));
        }
        if (lane == 0) {
            sx = __fadd_rn(sx, 13.0f); sy = __fadd_rn(sy, 13.0f); sz = __fadd_rn(sz, 13.0f);
            const size_t pix = (size_t)py * P.W + px;
            P.rgba[pix] = pack_rgba8_rz(sx, sy, sz, 255.0f);
            if (P.accum) P.accum[pix] = make_float4(sx, sy, sz, 255.0f);
        }
    }
    flush_counters(P, cnt, S->ntri_counted, P.ap.nsq + P.ap.nsp);
}

template <int VARIANT, bool FMA, int MEM>
static int launch_sm_pixel(pt_ctx ctx, const LaunchArgs &args_in) {
    LaunchArgs args = args_in;
    args.ap.tri_coop = MEM == PT_SCENE_SMEM;
    const uint32_t tiles_x = (uint32_t)(args.W + 7) / 8, tiles_y = (uint32_t)(args.nrows + 3) / 4;
    const uint32_t nitems = tiles_x * tiles_y * 32u;
    args.scatter_mul = 0;
    if (VARIANT != PT_VARIANT_GRID && ctx->h_scene[0]->ntri > 0 && nitems > 4096) {
        static const uint32_t primes[] = {40503u, 48271u, 69621u, 16807u, 65537u, 104729u};
        for (uint32_t pr : primes)
            if (std::gcd(pr, nitems) == 1u) { args.scatter_mul = pr; break; }
    }
    const size_t smem = MEM == PT_SCENE_SMEM ? (size_t)args.scene_bytes : 0;
    auto kern = k_sm_pixel<VARIANT, FMA, MEM, true>;
    int per_sm = 0;
    PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem), "occupancy query");
    if (per_sm < 1) per_sm = 1;
    uint32_t blocks = (uint32_t)(ctx->sm_count * per_sm);
    const uint32_t need_blocks = (nitems + 127) / 128;
    if (blocks > need_blocks) blocks = need_blocks;
    if (pt_ensure_scratch(ctx, 256)) return 1;
    uint32_t *counter = (uint32_t *)ctx->d_scratch;
    const uint32_t first_free = blocks * 128u;       // items [0, first_free) are the initial assignment
    PT_CUDA(cudaMemcpyAsync(counter, &first_free, 4, cudaMemcpyHostToDevice, ctx->stream), "init work counter");
    kern<<<blocks, 128, smem, ctx->stream>>>(args, nitems, counter);
    PT_CUDA(cudaGetLastError(), "launch k_sm_pixel");
    return 0;
}

template <bool FMA, int MEM>
static int launch_sm_nodof(pt_ctx ctx, const LaunchArgs &args) {
    const uint32_t tiles_x = (uint32_t)(args.W + 3) / 4, tiles_y = (uint32_t)(args.nrows + 1) / 2;
    const uint32_t nitems = tiles_x * tiles_y * 8u;      // one item = one pixel = one warp-pass
    const size_t smem = MEM == PT_SCENE_SMEM ? (size_t)args.scene_bytes : 0;
    auto kern = k_sm_nodof<FMA, MEM>;
    int per_sm = 0;
    PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem), "occupancy query");
    if (per_sm < 1) per_sm = 1;
    uint32_t blocks = (uint32_t)(ctx->sm_count * per_sm);
    const uint32_t need_blocks = (nitems + 7) / 8;
    if (blocks > need_blocks) blocks = need_blocks;
    kern<<<blocks, 256, smem, ctx->stream>>>(args, nitems);
    PT_CUDA(cudaGetLastError(), "launch k_sm_nodof");
    return 0;
}

template <bool FMA, int MEM>
static int launch_sm_am(pt_ctx ctx, int variant, const LaunchArgs &args) {
    switch (variant) {
        case PT_VARIANT_BASE: return launch_sm_pixel<PT_VARIANT_BASE, FMA, MEM>(ctx, args);
        case PT_VARIANT_LMEM: return launch_sm_pixel<PT_VARIANT_LMEM, FMA, MEM>(ctx, args);
        case PT_VARIANT_GRID: return launch_sm_pixel<PT_VARIANT_GRID, FMA, MEM>(ctx, args);
        case PT_VARIANT_NODOF: return launch_sm_nodof<FMA, MEM>(ctx, args);
    }
    return pt_fail(1, "unknown variant %d", variant);
}

}  // namespace pt

int pt_launch_persistent(pt_ctx ctx, const pt_render_params *p, const pt::LaunchArgs &args) {
    using namespace pt;
    const bool fma = p->arith != PT_ARITH_SEPARATE;
    if (p->scene_mem == PT_SCENE_SMEM)
        return fma ? launch_sm_am<true, PT_SCENE_SMEM>(ctx, p->variant, args)
                   : launch_sm_am<false, PT_SCENE_SMEM>(ctx, p->variant, args);
    int rc = pt_bind_const_scene(ctx, fma ? PT_ARITH_FMA : PT_ARITH_SEPARATE);
    if (rc) return rc;
    return fma ? launch_sm_am<true, PT_SCENE_CONST>(ctx, p->variant, args)
               : launch_sm_am<false, PT_SCENE_CONST>(ctx, p->variant, args);
}
