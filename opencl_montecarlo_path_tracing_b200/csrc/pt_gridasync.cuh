// pt_gridasync.cuh — trianglegrid variant, PT_KERNEL_GRID_ASYNC: the lanes of a warp keep THEIR pixels but not the same pace.
//
// Why.  The megakernel's warp traces 32 rays in lock step and waits for the longest walk: on the 1 M-triangle soup a ray
// visits 29 cells on average, the longest of 32 about 72 — 40 % of the lanes are in the DDA loop (ncu, profiles/r2_19).
// PT_KERNEL_GRID_STREAM attacked that with persistent lanes that fetch new pixels from a global counter; it lost (refilled
// lanes hold unrelated rays, and its refill ran whenever FEW LANES WALKED — i.e. all the time, with 5 lanes, as soon as short
// shadow rays came and went).  Here:
//   * a warp is an 8x4 pixel tile for its whole life, as in the megakernel — the rays of its lanes stay neighbours in the grid
//     whatever sample each lane is at (every lane draws from its own pixel's RNG stream, so the order of the draws of a pixel,
//     and with it every bit of the result, is untouched: SURVEY 0.3);
//   * the unit of lock-step work is ONE CELL VISIT (DDA step + the cell's records, software-pipelined as trace_grid);
//   * a lane whose ray has ended WAITS until ASYNC_K lanes wait (or nobody walks); then those lanes together consume their
//     results (shade / next light / next sample), generate their next rays, run the analytic tests and enter the grid — one
//     pass of straight-line code at >= ASYNC_K of 32 lanes instead of once per lane.
// Dead shadow rays (AnalyticParams::elide_dead) are honoured, so on the soup most samples are a single camera ray.
#pragma once
#include "pt_mega.cuh"

namespace pt {

#define ASYNC_K 16        // refill when this many lanes wait for their next ray
#define ASYNC_BATCH 4     // cell visits between two looks at the refill condition

template <bool FMA>
__global__ void __launch_bounds__(128, 6) k_grid_async(const __grid_constant__ LaunchArgs P) {
    typedef Ar<FMA> A;
    const SceneBlock *S = &c_scene;
    const GridDev &G = P.grid;
    const AnalyticParams &AP = P.ap;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long t_start = 0;
    if (P.cta_times && threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_start));
    uint32_t bx = blockIdx.x, by = blockIdx.y;
    if (P.tile_order) {
        const uint32_t tile = __ldg(P.tile_order + blockIdx.x);
        by = tile / P.tiles_x; bx = tile - by * P.tiles_x;
    }
    const int pi = bx * 16 + (warp & 1) * 8 + (lane & 7);
    const int vr = by * 8 + (warp >> 1) * 4 + (lane >> 3);
    const int pj = map_row(P, vr);
    Counters cnt = {0, 0, 0, 0, 0, 0};
    bool alive = pi < P.W && vr < P.nrows && pj < P.row_end;      // this lane's pixel still has samples to render
    Rng rng = rng_seed(P.seeds, alive ? (uint32_t)(pj * P.W + pi) : 0u);
    float cx = P.c0, cy = P.c0, cz = P.c0;
    int s = 0;
    // sample state (Sample(), grid:203-283): l = -1 camera ray in flight, l >= 0 shadow ray towards light l in flight
    int l = -1, m = 0, hit = HIT_NONE;
    V3 n = mk3(0.f, 0.f, 1.f), o = mk3(0.f, 0.f, 0.f), d = mk3(0.f, 0.f, 1.f);     // o: ray origin = hit point X while l >= 0
    float illum = 0.f, lam = 0.f, matf = 0.f, t = 1e9f;
    // DDA cursor
    float n0 = 0.f, n1 = 0.f, n2 = 0.f, dl0 = 0.f, dl1 = 0.f, dl2 = 0.f;
    int lin = 0;
    uint2 cell = make_uint2(0u, 0u);
    bool walking = false;          // a grid walk is in flight
    bool pending = false;          // a ray has ended; (hit, t) wait to be consumed
    const int K = P.scatter_mul ? (int)(P.scatter_mul & 0xff) : ASYNC_K, batch = P.scatter_mul ? (int)(P.scatter_mul >> 8) : ASYNC_BATCH;
    const int sx = G.pad_sx, sxy = G.pad_sxy;
    for (;;) {
        // ------------------------------------------------------------------------------------------------ REFILL
        const unsigned wmask = __ballot_sync(0xffffffffu, walking);
        const unsigned need = __ballot_sync(0xffffffffu, alive && !walking);
        if (need && (__popc(need) >= K || !wmask)) {
#pragma unroll 1
            for (int rep = 0; rep < 4; ++rep) {
                if (alive && !walking) {
                    bool start = true;
                    if (pending) {                                   // ---- consume the finished ray
                        pending = false;
                        bool sample_done = false;
                        V3 c = mk3(0.f, 0.f, 0.f);
                        if (l < 0) {
                            cnt.samples++;
                            if (hit == HIT_NONE) { c = shade_sky<FMA>(d); sample_done = true; }
                            else {
                                m = hit_material(hit);
                                n = hit_normal<FMA, true>(AP, S, G, hit, o, d, t);
                                o = A::vmadd(d, t, o);               // X
                                illum = 0.0f;
                                matf = 0.0f;
                                if (m == 1) {                        // checkerboard parity (grid:259-262)
                                    float yx = A::mul(o.x, 0.2f), yy = A::mul(o.y, 0.2f);
                                    matf = (f2i_rz_sat(A::add(ceilf(yx), ceilf(yy))) & 1) ? 1.0f : 0.0f;
                                } else if (m == 4) {                 // facing ratio (grid:268-270)
                                    float fr = A::dot(n, mk3(-d.x, -d.y, -d.z));
                                    matf = 0.0f < fr ? fr : 0.0f;
                                }
                                l = 0;
                                if (AP.elide_dead && m == 4) {       // dead shadow rays: only their RNG pairs are drawn
                                    for (int k = 0; k < AP.nlights; ++k) rng_skip(rng);
                                    l = AP.nlights;
                                }
                            }
                        } else {
                            if (hit == HIT_NONE) illum = light_add<FMA>(AP.lights[l], o, lam, illum);
                            l++;
                        }
                        if (!sample_done) {
                            bool shadow = false;
                            while (l < AP.nlights) {                 // next light that needs a shadow ray (grid:232-247)
                                float r0, r1;
                                rng_next(rng, r0, r1);               // drawn before any skip
                                V3 ld; float lm;
                                light_dir<FMA>(AP.lights[l], r0, r1, o, n, ld, lm);
                                if (lm < 0.0f) { l++; continue; }
                                d = ld; lam = lm;
                                cnt.shadow++;
                                shadow = true;
                                break;
                            }
                            if (!shadow) {
                                float il = illum;
                                if (il > 1.0f) il = 1.0f;
                                il = A::mul(il, 0.25f);
                                if (m == 1) { float i3 = A::mul(3.0f, il); c = matf != 0.0f ? mk3(i3, il, il) : mk3(i3, i3, i3); }
                                else if (m == 3) { float i2 = A::mul(2.0f, il); c = mk3(i2, A::mul(3.0f, il), i2); }
                                else c = mk3(matf, matf, matf);
                                sample_done = true;
                            }
                        }
                        if (sample_done) {
                            l = -1;
                            cx = A::madd(c.x, P.scale, cx);
                            cy = A::madd(c.y, P.scale, cy);
                            cz = A::madd(c.z, P.scale, cz);
                            if (++s == P.spp) {
                                const size_t pix = (size_t)pj * P.W + pi;
                                P.rgba[pix] = pack_rgba8_rz(cx, cy, cz, P.alpha);
                                if (P.accum) P.accum[pix] = make_float4(cx, cy, cz, P.alpha);
                                if (P.rng_out) P.rng_out[pix] = make_uint4(rng.x0, rng.x1, rng.c0, rng.c1);
                                alive = false;
                                start = false;
                            }
                        }
                    }
                    if (start) {                                     // ---- the next ray: TraceRay up to the grid entry (grid:102-176)
                        if (l < 0) {
                            camera_ray<FMA>(P.cam, rng, pi, pj, o, d);
                            t = 1e9f;                                // grid:222
                        }
                        cnt.rays++;
                        hit = HIT_NONE;
                        trace_analytic<FMA, true>(AP, S, o, d, t, hit);
                        float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
                        float tE[3], tX[3];
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            float inv = A::rcp(dd[a]);
                            float l1 = A::mul(A::sub(G.bmin[a], oo[a]), inv);
                            float l2 = A::mul(A::sub(G.bmax[a], oo[a]), inv);
                            tE[a] = cl_fmin(l1, l2);
                            tX[a] = cl_fmax(l1, l2);
                        }
                        float t0 = cl_fmax(cl_fmax(tE[0], tE[1]), cl_fmax(tE[0], tE[2]));
                        float t1 = cl_fmin(cl_fmin(tX[0], tX[1]), cl_fmin(tX[0], tX[2]));
                        if (t0 > t1) pending = true;                 // the ray misses the box: TraceRay is over
                        else {
                            bool inside = o.x >= G.bmin[0] && o.x <= G.bmax[0] && o.y >= G.bmin[1] && o.y <= G.bmax[1] &&
                                          o.z >= G.bmin[2] && o.z <= G.bmax[2];
                            float next[3], dl[3];
                            int idx[3];
#pragma unroll
                            for (int a = 0; a < 3; ++a) {
                                float p = inside ? oo[a] : A::madd(dd[a], t0, oo[a]);
                                int hi = G.res[a] - 1;
                                int v = f2i_rz_sat(A::div(A::sub(p, G.bmin[a]), G.cell[a]));
                                idx[a] = min(max(v, 0), hi);
                                dl[a] = A::div(A::sub(tX[a], tE[a]), __int2float_rn(G.res[a]));
                                bool pos = dd[a] > 0.0f;
                                next[a] = A::madd(__int2float_rn(pos ? idx[a] + 1 : G.res[a] - idx[a]), dl[a], tE[a]);
                            }
                            n0 = next[0]; n1 = next[1]; n2 = next[2];
                            dl0 = dl[0]; dl1 = dl[1]; dl2 = dl[2];
                            lin = (idx[2] + 1) * sxy + (idx[1] + 1) * sx + (idx[0] + 1);
                            cell = __ldg(G.cells_pad + lin);
                            walking = true;
                        }
                    }
                }
                // rays that ended at once (missed the box) go round again if enough lanes are in that position
                const unsigned again = __ballot_sync(0xffffffffu, alive && !walking);
                if (!again || (__popc(again) < K && __any_sync(0xffffffffu, walking))) break;
            }
        }
        if (!__any_sync(0xffffffffu, walking)) {
            if (!__any_sync(0xffffffffu, alive)) break;
            continue;
        }
        // -------------------------------------------------------------------------------------------------- WALK
        for (int it = 0; it < batch; ++it) {
            if (walking) {
                // one cell visit: the step to the next cell first (its word is requested before this cell's records are tested)
                const bool p01 = n0 < n1, p02 = n0 < n2, p12 = n1 < n2;
                const bool a0 = p01 & p02, a1 = (!p01) & p12;
                float lim;
                if (a0)      { n0 = A::add(n0, dl0); lim = n0; lin += d.x > 0.0f ? 1 : -1; }
                else if (a1) { n1 = A::add(n1, dl1); lim = n1; lin += d.y > 0.0f ? sx : -sx; }
                else         { n2 = A::add(n2, dl2); lim = n2; lin += d.z > 0.0f ? sxy : -sxy; }
                const uint2 ncell = __ldg(G.cells_pad + lin);
                cnt.cells++;
                cnt.gtri += cell.y;
                const float4 *rec = G.recs + 3 * (size_t)cell.x;
                uint32_t kb = 0xFFFFFFFFu;
                for (uint32_t k = 0; k < cell.y; ++k, rec += 3) {
                    float4 ra = __ldg(rec), rb = __ldg(rec + 1), rc = __ldg(rec + 2);
                    if (tri_test<FMA>(ra, rb, rc, o, d, t)) kb = k;
                }
                if (kb != 0xFFFFFFFFu) hit = hit_make(HIT_TRI, (int)(cell.x + kb));
                if (t < lim || ncell.y == 0xFFFFFFFFu) { walking = false; pending = true; }   // t compared AFTER the increment
                cell = ncell;
            }
        }
    }
    cnt.btests = cnt.gtri;
    flush_counters(P, cnt, 0, AP.nsq + AP.nsp);
    if (P.cta_times && threadIdx.x == 0) {
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end));
        const size_t b = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
        P.cta_times[2 * b] = t_start; P.cta_times[2 * b + 1] = t_end;
    }
}

template <bool FMA>
static int launch_grid_async(pt_ctx ctx, const LaunchArgs &args_in) {
    LaunchArgs args = args_in;
    args.scatter_mul = 0;
    if (getenv("PT_ASYNC")) args.scatter_mul = (uint32_t)atoi(getenv("PT_ASYNC"));   // tuning sweep: K | batch << 8
    return launch_pixel_b<PT_VARIANT_GRID, FMA, PT_SCENE_CONST, true>(ctx, args, k_grid_async<FMA>);
}

}  // namespace pt

int pt_launch_grid_async(pt_ctx ctx, const pt_render_params *p, const pt::LaunchArgs &args) {
    using namespace pt;
    const bool fma = p->arith != PT_ARITH_SEPARATE;
    int rc = pt_bind_const_scene(ctx, fma ? PT_ARITH_FMA : PT_ARITH_SEPARATE);
    if (rc) return rc;
    return fma ? launch_grid_async<true>(ctx, args) : launch_grid_async<false>(ctx, args);
}
