// pt_gridpool.cuh — trianglegrid variant, PT_KERNEL_GRID_POOL: a POOL of POOL_M pixels per lane in shared memory.
//
// Why.  On the 1 M-triangle soup 88 % of the megakernel's warp instructions are the grid traversal, and a warp walks it
// with 13.6 of the 24.4 lanes that entered (ncu source view, profiles/r2_04): ray lengths are independent (mean 14.35
// cells), every lane waits for the longest of its warp.  Ray regeneration has to hand an idle lane a NEW ray without
// (a) running the expensive shade / generate / grid-entry code for a handful of lanes at a time and (b) keeping two
// rays' worth of state in registers — the two things that sank PT_KERNEL_GRID_STREAM (profiles/r2_05: its refill code
// ran with 3-5 lanes, 88 registers).  Here
//   * every lane owns POOL_M pixels ("slots"); the complete state of a slot — RNG stream, colour sum, sample state,
//     ray, DDA cursor: 34 words — lives in shared memory, lane-interleaved (bank = lane: conflict-free whichever slot a
//     lane picks), so the register budget is that of ONE phase;
//   * the warp votes between two phases.  TRAVERSE: every lane that has a slot with a live traversal loads it, visits
//     POOL_Q cells (sphere prefilter + Moller-Trumbore as trace_grid), stores it back.  SHADE: every lane that has a
//     slot whose ray ended consumes the result (Sample(), grid:203-283), starts the slot's next ray — shadow ray, next
//     sample, or a new pixel fetched with one warp-aggregated atomicAdd — runs the analytic tests and the grid entry.
//     A lane whose one slot waits for the SHADE phase keeps traversing with its other slot, so both phases run with
//     most of the warp;
//   * work items are handed out tile-major (8x4 pixels), so the pixels a warp holds stay close to each other.
// Per ray the operations and their order are those of TraceRay / Sample (grid:102-283); only which rays share a warp at
// a given moment changes: results (image, float sums, RNG states, counters) stay bit-identical.
#pragma once
#include "pt_persistent.cuh"

namespace pt {

#define POOL_M 2          // slots (pixels) per lane
#define POOL_WORDS 34     // 32-bit words of state per slot

enum { PS_EMPTY = 0, PS_SHADE = 1, PS_TRAV = 2, PS_DEAD = 3 };

PT_DEV int pool_state(uint32_t st, int s) { return (st >> (2 * s)) & 3; }
PT_DEV uint32_t pool_set(uint32_t st, int s, int v) { return (st & ~(3u << (2 * s))) | ((uint32_t)v << (2 * s)); }

// slab test + DDA initialisation (grid:157-176), the part of trace_grid before its loop; false: the ray misses the box
template <bool FMA>
PT_DEV bool grid_enter(const GridDev &G, V3 o, V3 d, float next[3], float dl[3], int idx[3], uint2 &cell) {
    typedef Ar<FMA> A;
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
    if (t0 > t1) return false;
    bool inside = o.x >= G.bmin[0] && o.x <= G.bmax[0] && o.y >= G.bmin[1] && o.y <= G.bmax[1] &&
                  o.z >= G.bmin[2] && o.z <= G.bmax[2];
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
    cell = __ldg(&G.cells[(size_t)idx[2] * (G.res[0] * G.res[1]) + (size_t)idx[1] * G.res[0] + idx[0]]);
    return true;
}

// One iteration of trace_grid's loop (grid:177-198): step to the next cell, request its word, test the current cell's
// triangles, the reference's termination test.  true: the traversal goes on (cell = the next cell's word).
// The cursor is kept in a form that needs no per-axis branches (lanes of a pool warp step along different axes):
//   ci  = linear index of the cell the NEXT step leads to ... after the step; rem = steps left before the ray leaves the
//   grid, 10 bits per axis; the per-axis linear strides are recomputed from the direction's signs.
// `mask` = the lanes that execute this call together (they re-converge behind each divergent part).
template <bool FMA>
PT_DEV bool grid_visit(const GridDev &G, unsigned mask, V3 o, V3 d, float &t, int &hit, float &n0, float &n1, float &n2, float d0, float d1,
                       float d2, int &ci, uint32_t &rem, int s0, int s1, int s2, uint2 &cell, Counters &cnt) {
    typedef Ar<FMA> A;
    const int kk = ((n0 < n1) << 2) + ((n0 < n2) << 1) + (n1 < n2);
    const int axis = (0x00221212u >> (4 * kk)) & 0xF;          // the reference's LUT {2,1,2,1,2,2,0,0}
    const bool a0 = axis == 0, a1 = axis == 1;
    const float m0 = A::add(n0, d0), m1 = A::add(n1, d1), m2 = A::add(n2, d2);
    n0 = a0 ? m0 : n0;
    n1 = a1 ? m1 : n1;
    n2 = (a0 || a1) ? n2 : m2;
    const float lim = a0 ? m0 : (a1 ? m1 : m2);
    const int sh = 10 * axis;
    const bool at_end = ((rem >> sh) & 1023u) == 0u;            // no step left along this axis: it leaves the grid
    rem -= 1u << sh;
    ci += a0 ? s0 : (a1 ? s1 : s2);
    uint2 ncell = make_uint2(0u, 0u);
    if (!at_end) ncell = __ldg(&G.cells[ci]);
    cnt.cells++;
    cnt.gtri += cell.y;
    for (uint32_t base = 0; base < cell.y; base += 32) {
        const uint32_t n = min(cell.y - base, 32u);
        const float4 *sp = G.sph + (size_t)cell.x + base;
        uint32_t surv = 0;
        for (uint32_t k = 0; k < n; ++k)
            if (line_near_sphere(__ldg(sp + k), G.sph_k, o, d)) surv |= 1u << k;
        cnt.btests += __popc(surv);
        while (surv) {
            const uint32_t k = base + (uint32_t)__ffs((int)surv) - 1u;
            surv &= surv - 1u;
            const float4 *rec = G.recs + 3 * ((size_t)cell.x + k);
            float4 ra = __ldg(rec), rb = __ldg(rec + 1), rc = __ldg(rec + 2);
            if (tri_test<FMA>(ra, rb, rc, o, d, t)) hit = hit_make(HIT_TRI, (int)(cell.x + k));
        }
    }
    __syncwarp(mask);                                           // cells hold different numbers of records / survivors
    const bool go = !(t < lim || at_end);                       // t compared AFTER the increment (grid:194-195)
    cell = go ? ncell : cell;
    return go;
}

template <bool FMA>
__global__ void __launch_bounds__(128, 6) k_grid_pool(const __grid_constant__ LaunchArgs P, uint32_t nitems, uint32_t *work_counter,
                                                      int thresh_shade, int quantum) {
    typedef Ar<FMA> A;
    extern __shared__ __align__(16) uint32_t pool_smem[];
    const SceneBlock *S = &c_scene;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *wsm = pool_smem + warp * (POOL_WORDS * POOL_M * 32) + lane;
#define PW(w, s) wsm[((w) * POOL_M + (s)) * 32]
#define PF(w, s) __uint_as_float(PW(w, s))
#define PSTF(w, s, v) PW(w, s) = __float_as_uint(v)
    Counters cnt = {0, 0, 0, 0, 0, 0};
    uint32_t st = 0;                                           // 2 bits per slot: all PS_EMPTY
    const unsigned lt_mask = (1u << lane) - 1u;
    for (;;) {
        // ------------------------------------------------------------------ pixels for the empty slots
#pragma unroll
        for (int s = 0; s < POOL_M; ++s) {
            const bool want = pool_state(st, s) == PS_EMPTY;
            const unsigned need = __ballot_sync(0xffffffffu, want);
            if (need) {                                        // warp-uniform
                const int leader = __ffs(need) - 1;
                uint32_t base = 0;
                if ((int)lane == leader) base = atomicAdd(work_counter, (uint32_t)__popc(need));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (want) {
                    const uint32_t w = base + __popc(need & lt_mask);
                    int i, j;
                    if (w >= nitems) {
                        st = pool_set(st, s, PS_DEAD);         // queue exhausted
                    } else if (item_to_pixel(P, w, i, j)) {
                        const Rng r = rng_seed(P.seeds, (uint32_t)(j * P.W + i));
                        PW(0, s) = r.x0; PW(1, s) = r.x1; PW(2, s) = r.c0; PW(3, s) = r.c1;
                        PSTF(4, s, P.c0); PSTF(5, s, P.c0); PSTF(6, s, P.c0);
                        PW(7, s) = 0u;                         // samples done
                        PW(8, s) = (uint32_t)i | ((uint32_t)j << 16);
                        PW(22, s) = 1u << 8;                   // phase 0, fresh: no result to consume
                        st = pool_set(st, s, PS_SHADE);
                    }                                          // else: the item lies outside the image, draw another next round
                }
            }
        }
        bool hasT = false, hasS = false, hasE = false;
#pragma unroll
        for (int s = 0; s < POOL_M; ++s) {
            const int v = pool_state(st, s);
            hasT |= v == PS_TRAV; hasS |= v == PS_SHADE; hasE |= v == PS_EMPTY;
        }
        const unsigned mT = __ballot_sync(0xffffffffu, hasT), mS = __ballot_sync(0xffffffffu, hasS);
        if (!(mT | mS)) {
            if (!__any_sync(0xffffffffu, hasE)) break;         // every slot of the warp is dead: done
            continue;
        }
        if (__popc(mS) >= thresh_shade || !mT) {
            // -------------------------------------------------------------- SHADE: consume a finished ray, start the next
            if (hasS) {
                int s = 0;
#pragma unroll
                for (int q = POOL_M - 1; q >= 0; --q) if (pool_state(st, q) == PS_SHADE) s = q;
                Lane L;
                L.rng.x0 = PW(0, s); L.rng.x1 = PW(1, s); L.rng.c0 = PW(2, s); L.rng.c1 = PW(3, s);
                float cx = PF(4, s), cy = PF(5, s), cz = PF(6, s);
                int sdone = (int)PW(7, s);
                const uint32_t pxy = PW(8, s);
                L.px = (int)(pxy & 0xffffu); L.py = (int)(pxy >> 16);
                L.o = mk3(PF(9, s), PF(10, s), PF(11, s));
                L.X = L.o;                                     // a shadow ray starts at X: one copy serves both
                L.d = mk3(PF(12, s), PF(13, s), PF(14, s));
                L.t = PF(15, s);
                int hit = (int)PW(16, s);
                L.n = mk3(PF(17, s), PF(18, s), PF(19, s));
                L.illum = PF(20, s); L.lam = PF(21, s);
                const uint32_t fl = PW(22, s);
                L.phase = fl & 3; L.l = (fl >> 2) & 7; L.mat = (fl >> 5) & 7;
                L.matf = PF(23, s);
                bool start = true;
                if (!((fl >> 8) & 1u)) {                       // ---- a result waits (Sample, grid:203-283)
                    bool sample_done = false;
                    V3 c = mk3(0.f, 0.f, 0.f);
                    if (L.phase == 0) {
                        cnt.samples++;
                        if (hit == HIT_NONE) { c = shade_sky<FMA>(L.d); sample_done = true; }
                        else {
                            L.mat = hit_material(hit);
                            L.n = hit_normal<FMA, true>(P.ap, S, P.grid, hit, L.o, L.d, L.t);
                            L.X = A::vmadd(L.d, L.t, L.o);
                            L.illum = 0.0f;
                            L.matf = 0.0f;
                            if (L.mat == 1) {
                                float yx = A::mul(L.X.x, 0.2f), yy = A::mul(L.X.y, 0.2f);
                                L.matf = (f2i_rz_sat(A::add(ceilf(yx), ceilf(yy))) & 1) ? 1.0f : 0.0f;
                            } else if (L.mat == 4) {
                                float fr = A::dot(L.n, mk3(-L.d.x, -L.d.y, -L.d.z));
                                L.matf = 0.0f < fr ? fr : 0.0f;
                            }
                            L.l = 0;
                        }
                    } else {
                        if (hit == HIT_NONE) L.illum = light_add<FMA>(P.ap.lights[L.l], L.X, L.lam, L.illum);
                        L.l++;
                    }
                    if (!sample_done && !next_shadow_ray<FMA, true>(P.ap, L, cnt)) {
                        c = finish_material<FMA>(L);
                        sample_done = true;
                    }
                    if (sample_done) {
                        L.phase = 0;
                        cx = A::madd(c.x, P.scale, cx);
                        cy = A::madd(c.y, P.scale, cy);
                        cz = A::madd(c.z, P.scale, cz);
                        if (++sdone == P.spp) {
                            const size_t pix = (size_t)L.py * P.W + L.px;
                            P.rgba[pix] = pack_rgba8_rz(cx, cy, cz, P.alpha);
                            if (P.accum) P.accum[pix] = make_float4(cx, cy, cz, P.alpha);
                            if (P.rng_out) P.rng_out[pix] = make_uint4(L.rng.x0, L.rng.x1, L.rng.c0, L.rng.c1);
                            st = pool_set(st, s, PS_EMPTY);
                            start = false;
                        }
                    }
                }
                __syncwarp(mS);                                // the paths through Sample() above re-converge here
                if (start) {                                   // ---- begin the next ray (TraceRay up to the grid entry, grid:102-176)
                    if (L.phase == 0) {
                        camera_ray<FMA>(P.cam, L.rng, L.px, L.py, L.o, L.d);
                        L.t = 1e9f;                            // grid:222
                    }
                    cnt.rays++;
                    hit = HIT_NONE;
                    trace_analytic<FMA, true>(P.ap, S, L.o, L.d, L.t, hit);
                    float next[3], dl[3];
                    int idx[3];
                    uint2 cell;
                    const bool trav = grid_enter<FMA>(P.grid, L.o, L.d, next, dl, idx, cell);
                    PW(0, s) = L.rng.x0; PW(1, s) = L.rng.x1; PW(2, s) = L.rng.c0; PW(3, s) = L.rng.c1;
                    PSTF(4, s, cx); PSTF(5, s, cy); PSTF(6, s, cz);
                    PW(7, s) = (uint32_t)sdone;
                    PSTF(9, s, L.o.x); PSTF(10, s, L.o.y); PSTF(11, s, L.o.z);
                    PSTF(12, s, L.d.x); PSTF(13, s, L.d.y); PSTF(14, s, L.d.z);
                    PSTF(15, s, L.t);
                    PW(16, s) = (uint32_t)hit;
                    PSTF(17, s, L.n.x); PSTF(18, s, L.n.y); PSTF(19, s, L.n.z);
                    PSTF(20, s, L.illum); PSTF(21, s, L.lam);
                    PW(22, s) = (uint32_t)L.phase | ((uint32_t)L.l << 2) | ((uint32_t)L.mat << 5);   // a result will wait
                    PSTF(23, s, L.matf);
                    if (trav) {
                        PSTF(24, s, next[0]); PSTF(25, s, next[1]); PSTF(26, s, next[2]);
                        PSTF(27, s, dl[0]); PSTF(28, s, dl[1]); PSTF(29, s, dl[2]);
                        // steps left before the ray leaves the grid, per axis, and the linear index of its first cell
                        const uint32_t r0 = (uint32_t)(L.d.x > 0.0f ? P.grid.res[0] - 1 - idx[0] : idx[0]);
                        const uint32_t r1 = (uint32_t)(L.d.y > 0.0f ? P.grid.res[1] - 1 - idx[1] : idx[1]);
                        const uint32_t r2 = (uint32_t)(L.d.z > 0.0f ? P.grid.res[2] - 1 - idx[2] : idx[2]);
                        PW(30, s) = r0 | (r1 << 10) | (r2 << 20);
                        PW(31, s) = (uint32_t)((idx[2] * P.grid.res[1] + idx[1]) * P.grid.res[0] + idx[0]);
                        PW(32, s) = cell.x; PW(33, s) = cell.y;
                        st = pool_set(st, s, PS_TRAV);
                    }                                          // else: the ray misses the grid box; its result waits for the next SHADE
                }
            }
        } else {
            // -------------------------------------------------------------- TRAVERSE: `quantum` cell visits of one live slot
            int s = 0;
#pragma unroll
            for (int q = POOL_M - 1; q >= 0; --q) if (pool_state(st, q) == PS_TRAV) s = q;
            bool act = hasT;
            V3 o = mk3(0.f, 0.f, 0.f), d = o;
            float t = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f;
            int hit = 0, ci = 0, s0 = 0, s1 = 0, s2 = 0;
            uint32_t rem = 0;
            uint2 cell = make_uint2(0u, 0u);
            if (act) {
                o = mk3(PF(9, s), PF(10, s), PF(11, s));
                d = mk3(PF(12, s), PF(13, s), PF(14, s));
                t = PF(15, s);
                hit = (int)PW(16, s);
                n0 = PF(24, s); n1 = PF(25, s); n2 = PF(26, s);
                d0 = PF(27, s); d1 = PF(28, s); d2 = PF(29, s);
                rem = PW(30, s);
                ci = (int)PW(31, s);
                cell = make_uint2(PW(32, s), PW(33, s));
                s0 = d.x > 0.0f ? 1 : -1;
                s1 = d.y > 0.0f ? P.grid.res[0] : -P.grid.res[0];
                s2 = d.z > 0.0f ? P.grid.res[0] * P.grid.res[1] : -(P.grid.res[0] * P.grid.res[1]);
            }
            bool fin = false;
            for (int q = 0; q < quantum; ++q) {
                const unsigned am = __ballot_sync(0xffffffffu, act);
                if (!am) break;
                if (act) {
                    act = grid_visit<FMA>(P.grid, am, o, d, t, hit, n0, n1, n2, d0, d1, d2, ci, rem, s0, s1, s2, cell, cnt);
                    fin = !act;
                }
                __syncwarp();
            }
            if (hasT) {
                PSTF(15, s, t);
                PW(16, s) = (uint32_t)hit;
                if (fin) {
                    st = pool_set(st, s, PS_SHADE);            // the result (t, hit) waits for the SHADE phase
                } else {
                    PSTF(24, s, n0); PSTF(25, s, n1); PSTF(26, s, n2);
                    PW(30, s) = rem;
                    PW(31, s) = (uint32_t)ci;
                    PW(32, s) = cell.x; PW(33, s) = cell.y;
                }
            }
        }
        __syncwarp();
    }
#undef PW
#undef PF
#undef PSTF
    flush_counters(P, cnt, 0, P.ap.nsq + P.ap.nsp);
}

template <bool FMA>
static int launch_grid_pool(pt_ctx ctx, const LaunchArgs &args) {
    const uint32_t tiles_x = (uint32_t)(args.W + 7) / 8, tiles_y = (uint32_t)(args.nrows + 3) / 4;
    const uint32_t nitems = tiles_x * tiles_y * 32u;
    const size_t smem = (size_t)POOL_WORDS * POOL_M * 128 * 4;
    auto kern = k_grid_pool<FMA>;
    int per_sm = 0;
    PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem), "occupancy query");
    if (per_sm < 1) per_sm = 1;
    uint32_t blocks = (uint32_t)(ctx->sm_count * per_sm);
    const uint32_t need_blocks = (nitems + 128 * POOL_M - 1) / (128 * POOL_M);
    if (blocks > need_blocks) blocks = need_blocks;
    if (pt_ensure_scratch(ctx, 256)) return 1;
    uint32_t *counter = (uint32_t *)ctx->d_scratch;
    PT_CUDA(cudaMemsetAsync(counter, 0, 4, ctx->stream), "init work counter");
    int thresh = 16, quantum = 4;
    if (const char *e = getenv("PT_POOL")) { const int v = atoi(e); thresh = v & 0xff; quantum = (v >> 8) & 0xff; if (thresh < 1) thresh = 1; if (quantum < 1) quantum = 1; }
    kern<<<blocks, 128, smem, ctx->stream>>>(args, nitems, counter, thresh, quantum);
    PT_CUDA(cudaGetLastError(), "launch k_grid_pool");
    return 0;
}

}  // namespace pt

// eligible: 10-bit cell indices, 16-bit pixel coordinates
static bool pt_grid_pool_ok(const pt::LaunchArgs &args) {
    return args.grid.res[0] <= 1023 && args.grid.res[1] <= 1023 && args.grid.res[2] <= 1023 && args.W <= 65535 && args.H <= 65535;
}

int pt_launch_grid_pool(pt_ctx ctx, const pt_render_params *p, const pt::LaunchArgs &args) {
    using namespace pt;
    if (!pt_grid_pool_ok(args)) return pt_fail(1, "PT_KERNEL_GRID_POOL: grid resolution above 1023 or image side above 65535");
    const bool fma = p->arith != PT_ARITH_SEPARATE;
    int rc = pt_bind_const_scene(ctx, fma ? PT_ARITH_FMA : PT_ARITH_SEPARATE);
    if (rc) return rc;
    return fma ? launch_grid_pool<true>(ctx, args) : launch_grid_pool<false>(ctx, args);
}
