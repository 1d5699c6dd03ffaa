// pt_gridqueue.cuh — trianglegrid variant, PT_KERNEL_GRID_QUEUE: the lanes WALK the grid, the WARP tests the triangles.
//
// Why.  In the megakernel every lane walks its own DDA and tests the triangles of its own cells.  On the 1 M-triangle soup
// the walk runs with ~13 of 32 lanes (the others' rays have ended: a warp waits for its longest walk), and inside a step the
// warp pays the triangle loop of its FULLEST cell — 3-4 iterations for an average of 1.75 records per lane — and, per
// iteration, every stage of Moller-Trumbore as soon as ONE lane survives the stage before (24 % pass the first barycentric
// test: with ten lanes that is nearly always).  ncu, profiles/r2_19: 63 % of the thread instructions are triangle tests,
// issued with 8-10 lanes.
// Here a step is split in two:
//   * WALK (the lanes with a live ray): axis choice, boundary update, next cell word — as trace_grid, on the padded array;
//   * TEST (all 32 lanes, whether they own a ray or not): the walking lanes' (ray, record) pairs of this step are written
//     to a 32-entry queue in shared memory — a warp prefix sum of the cell counts gives every lane its slots — and each
//     lane takes ONE pair: record from global memory, ray (origin, direction, distance bound at cell entry) from the owner's
//     slot in shared memory, the same tri_test.  ~23 pairs per step fill one batch; longer lists take further batches.
//   * An accepted pair enters a per-owner 64-bit atomicMin on (distance, record index): the reference scans a cell in order
//     and keeps a hit only if it is STRICTLY closer (grid:61-85, 185-198), so after the cell t is the smallest accepted
//     distance below the entry bound and the hit is the first record that attains it — the lexicographic minimum.  -0 and +0
//     are one distance for that comparison (key from r + 0.0f); the winner's own bits (sign of zero included) become t.
// Per ray the operations are those of trace_grid / Sample: image, accumulation buffer, RNG states and counters stay
// bit-identical.  Sample() is laid out in warp-uniform ROUNDS (camera rays, then the shadow rays towards light 0, 1, ...) so
// that the whole warp is present at every walk, as in PT_KERNEL_SPEC's sample_rounds.
#pragma once
#include "pt_mega.cuh"

namespace pt {

#define GQ_OWNER_SHIFT 27u                      // queue word: owner lane << 27 | record index  (records < 2^27)

struct __align__(16) GqWarp {                    // shared memory of one warp: 1.5 KB
    float4 ray_o[32];                            // origin.xyz, distance bound t at the entry of the current cell
    float4 ray_d[32];                            // direction.xyz, -
    unsigned long long key[32];                  // per owner: min over accepted pairs of ordered(r + 0) << 32 | record index
    uint32_t rbits[32];                          // the winner's own r
    uint32_t queue[32];
};

// One TraceRay through the grid for the whole warp (grid:157-198).  `has_ray`: this lane traces (o, d) with bound t.
template <bool FMA>
PT_DEV void trace_grid_queue(const GridDev &G, GqWarp &W, const unsigned lane, bool has_ray, V3 o, V3 d, float &t, int &hit, Counters &cnt) {
    typedef Ar<FMA> A;
    // ---- slab test + DDA initialisation: the first half of trace_grid, per lane
    bool walking = false;
    float n0 = 0.f, n1 = 0.f, n2 = 0.f, dl0 = 0.f, dl1 = 0.f, dl2 = 0.f;
    int lin = 0;
    if (has_ray) {
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
        if (!(t0 > t1)) {
            walking = true;
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
            lin = (idx[2] + 1) * G.pad_sxy + (idx[1] + 1) * G.pad_sx + (idx[0] + 1);
            W.ray_o[lane] = make_float4(o.x, o.y, o.z, t);
            W.ray_d[lane] = make_float4(d.x, d.y, d.z, 0.f);
        }
    }
    if (!__any_sync(0xffffffffu, walking)) return;
    const bool pos0 = d.x > 0.0f, pos1 = d.y > 0.0f, pos2 = d.z > 0.0f;
    const int sx = G.pad_sx, sxy = G.pad_sxy;
    uint2 cell = make_uint2(0u, 0u);
    if (walking) cell = __ldg(G.cells_pad + lin);
    for (;;) {
        // ---- WALK: the step to the next cell and the request of its word (as trace_grid)
        float lim = 0.f;
        uint2 ncell = make_uint2(0u, 0u);
        uint32_t rem = 0u;
        if (walking) {
            const bool p01 = n0 < n1, p02 = n0 < n2, p12 = n1 < n2;
            const bool a0 = p01 & p02, a1 = (!p01) & p12;
            if (a0)      { n0 = A::add(n0, dl0); lim = n0; lin += pos0 ? 1 : -1; }
            else if (a1) { n1 = A::add(n1, dl1); lim = n1; lin += pos1 ? sx : -sx; }
            else         { n2 = A::add(n2, dl2); lim = n2; lin += pos2 ? sxy : -sxy; }
            ncell = __ldg(G.cells_pad + lin);
            cnt.cells++;
            cnt.gtri += cell.y;
            rem = cell.y;
            W.key[lane] = ~0ull;
        }
        // ---- TEST: batches of up to 32 (ray, record) pairs, one per lane
        uint32_t first = cell.x;
        bool any_hit = false;
        while (__any_sync(0xffffffffu, rem != 0u)) {
            // inclusive prefix sum of the remaining counts
            uint32_t incl = rem;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
                if ((int)lane >= off) incl += v;
            }
            const uint32_t excl = incl - rem;
            const uint32_t total = min(32u, __shfl_sync(0xffffffffu, incl, 31));
            const uint32_t take = excl >= 32u ? 0u : min(rem, 32u - excl);
            for (uint32_t q = 0; q < take; ++q) W.queue[excl + q] = (lane << GQ_OWNER_SHIFT) | (first + q);
            first += take;
            rem -= take;
            __syncwarp();
            bool acc = false;
            float r = 0.f;
            uint32_t owner = 0u, ri = 0u;
            if (lane < total) {
                const uint32_t w = W.queue[lane];
                owner = w >> GQ_OWNER_SHIFT; ri = w & ((1u << GQ_OWNER_SHIFT) - 1u);
                const float4 *rec = G.recs + 3 * (size_t)ri;
                const float4 ra = __ldg(rec), rb = __ldg(rec + 1), rc = __ldg(rec + 2);
                const float4 qo = W.ray_o[owner], qd = W.ray_d[owner];
                r = qo.w;
                acc = tri_test<FMA>(ra, rb, rc, mk3(qo.x, qo.y, qo.z), mk3(qd.x, qd.y, qd.z), r);
                if (acc) atomicMin(&W.key[owner], ((unsigned long long)ordered_key(r + 0.0f) << 32) | ri);
            }
            if (__any_sync(0xffffffffu, acc)) {
                any_hit = true;
                __syncwarp();                              // every atomicMin of the batch has landed
                if (acc && (uint32_t)W.key[owner] == ri) W.rbits[owner] = __float_as_uint(r);
            }
            __syncwarp();                                  // this batch's queue reads before the next batch's writes
        }
        if (any_hit) __syncwarp();
        // ---- the owners collect their result and decide whether the walk goes on
        if (walking) {
            if (any_hit) {
                const unsigned long long k = W.key[lane];
                if (k != ~0ull) {
                    t = __uint_as_float(W.rbits[lane]);
                    hit = hit_make(HIT_TRI, (int)(uint32_t)k);
                    W.ray_o[lane].w = t;
                }
            }
            if (t < lim || ncell.y == 0xFFFFFFFFu) walking = false;     // t compared AFTER the increment (grid:194-195)
            cell = ncell;
        }
        if (!__any_sync(0xffffffffu, walking)) break;
    }
}

// Sample() in warp-uniform rounds: round -1 the camera rays, round l the shadow rays towards light l (lanes without one idle
// but help with the triangle tests).  Per lane the operations and the order of its RNG draws are those of sample().
template <bool FMA>
PT_DEV V3 sample_rounds_grid(const AnalyticParams &AP, const SceneBlock *S, const GridDev &G, GqWarp &W, const unsigned lane, bool active, V3 o, V3 d,
                             Rng &rng, Counters &cnt) {
    typedef Ar<FMA> A;
    if (active) cnt.samples++;
    float t = 1e9f, illum = 0.0f, lam = 0.0f;
    V3 ro = o, rd = d, X = o, n = o, out = mk3(0.f, 0.f, 0.f);
    int m = 0;
    bool alive = active, has_ray = active;
    for (int l = -1; l < AP.nlights; ++l) {
        if (l >= 0) {
            has_ray = false;
            if (alive) {
                float r0, r1;
                rng_next(rng, r0, r1);                                  // drawn before any skip (grid:233)
                if (!(AP.elide_dead && m == 4)) {                       // dead shadow rays: AnalyticParams::elide_dead
                    light_dir<FMA>(AP.lights[l], r0, r1, X, n, rd, lam);
                    if (!(lam < 0.0f)) { ro = X; cnt.shadow++; has_ray = true; }
                }
            }
            if (!__any_sync(0xffffffffu, has_ray)) continue;
        }
        int hit = HIT_NONE;
        if (has_ray) {
            cnt.rays++;
            trace_analytic<FMA, true>(AP, S, ro, rd, t, hit);
        }
        trace_grid_queue<FMA>(G, W, lane, has_ray, ro, rd, t, hit, cnt);
        if (l < 0) {
            if (active) {
                if (hit == HIT_NONE) { out = shade_sky<FMA>(d); alive = false; }
                else {
                    m = hit_material(hit);
                    n = hit_normal<FMA, true>(AP, S, G, hit, o, d, t);
                    X = A::vmadd(d, t, o);
                }
            }
        } else if (has_ray && hit == HIT_NONE) {
            illum = light_add<FMA>(AP.lights[l], X, lam, illum);
        }
    }
    if (alive) out = shade_material<FMA>(m, illum, X, n, d);
    return out;
}

template <bool FMA>
__global__ void __launch_bounds__(128, 6) k_grid_queue(const __grid_constant__ LaunchArgs P) {
    __shared__ GqWarp s_warp[4];
    const SceneBlock *S = &c_scene;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long t_start = 0;
    if (P.cta_times && threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_start));
    uint32_t bx = blockIdx.x, by = blockIdx.y;
    if (P.tile_order) {
        const uint32_t tile = __ldg(P.tile_order + blockIdx.x);
        by = tile / P.tiles_x; bx = tile - by * P.tiles_x;
    }
    const int i = bx * 16 + (warp & 1) * 8 + (lane & 7);
    const int vr = by * 8 + (warp >> 1) * 4 + (lane >> 3);
    Counters cnt = {0, 0, 0, 0, 0, 0};
    const int j = map_row(P, vr);
    const bool mine = i < P.W && vr < P.nrows && j < P.row_end;
    Rng rng = rng_seed(P.seeds, mine ? (uint32_t)(j * P.W + i) : 0u);
    float cx = P.c0, cy = P.c0, cz = P.c0;
    if (__any_sync(0xffffffffu, mine)) {
        for (int s = 0; s < P.spp; ++s) {
            V3 o = mk3(0.f, 0.f, 0.f), d = mk3(0.f, 0.f, 1.f);
            if (mine) camera_ray<FMA>(P.cam, rng, i, j, o, d);
            const V3 c = sample_rounds_grid<FMA>(P.ap, S, P.grid, s_warp[warp], (unsigned)lane, mine, o, d, rng, cnt);
            cx = Ar<FMA>::madd(c.x, P.scale, cx);
            cy = Ar<FMA>::madd(c.y, P.scale, cy);
            cz = Ar<FMA>::madd(c.z, P.scale, cz);
        }
    }
    if (mine) {
        const size_t pix = (size_t)j * P.W + i;
        P.rgba[pix] = pack_rgba8_rz(cx, cy, cz, P.alpha);
        if (P.accum) P.accum[pix] = make_float4(cx, cy, cz, P.alpha);
        if (P.rng_out) P.rng_out[pix] = make_uint4(rng.x0, rng.x1, rng.c0, rng.c1);
    }
    cnt.btests = cnt.gtri;        // every record of a visited cell is tested
    flush_counters(P, cnt, 0, P.ap.nsq + P.ap.nsp);
    if (P.cta_times && threadIdx.x == 0) {
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end));
        const size_t b = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
        P.cta_times[2 * b] = t_start; P.cta_times[2 * b + 1] = t_end;
    }
}

template <bool FMA>
static int launch_grid_queue(pt_ctx ctx, const LaunchArgs &args) {
    return launch_pixel_b<PT_VARIANT_GRID, FMA, PT_SCENE_CONST, true>(ctx, args, k_grid_queue<FMA>);
}

}  // namespace pt

int pt_launch_grid_queue(pt_ctx ctx, const pt_render_params *p, const pt::LaunchArgs &args) {
    using namespace pt;
    if (ctx->total_refs >= (1ull << GQ_OWNER_SHIFT)) return pt_fail(1, "PT_KERNEL_GRID_QUEUE: the grid holds 2^27 or more records");
    const bool fma = p->arith != PT_ARITH_SEPARATE;
    int rc = pt_bind_const_scene(ctx, fma ? PT_ARITH_FMA : PT_ARITH_SEPARATE);
    if (rc) return rc;
    return fma ? launch_grid_queue<true>(ctx, args) : launch_grid_queue<false>(ctx, args);
}
