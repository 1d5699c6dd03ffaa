// pt_gridtma.cuh — warp-cooperative grid traversal with TMA-staged cell lists (PT_KERNEL_GRID_TMA).
//
// One WARP per ray.  The DDA state is warp-uniform; for every non-empty cell ONE elected lane issues a single
// bulk asynchronous copy (cp.async.bulk.shared::cluster.global -> SASS UBLKCP, the non-tensor TMA path) of the
// cell's contiguous triangle-record block (count x 48 B, <= 2976 B) into a shared-memory stage and arms an
// mbarrier with the byte count; the copy of the NEXT cell is issued before the current cell is tested
// (double buffering; the DDA axis choice does not depend on t, only the termination test does).  The 32
// lanes then test 32 different triangles of the staged list and a (t, index) shuffle reduction with
// "lowest index wins ties" reproduces the reference's first-found-wins scan order (grid:87-95).
//
// This is the design the task names for dense cells (up to 62 references).  For sparse cells — the default
// CELL_SIZE_MODIFIER gives 2-8 references per cell — a warp per ray wastes most lanes, which is why
// PT_KERNEL_AUTO keeps the thread-per-ray traversal; DESIGN.md reports both measured.
#pragma once
#include "pt_host.h"

namespace pt {

#define PT_TMA_STAGE_BYTES (62 * 48)

struct CoopStage {
    float4 *buf[2];      // two stages of 62 records
    uint64_t *bar[2];
    uint32_t parity[2];  // phase parity to wait for next, per stage (warp-uniform)
};

// lane 0 arms the barrier and issues the bulk copy of one cell's record block into stage s
PT_DEV void coop_issue(const GridDev &G, CoopStage &st, int s, uint2 cell, int lane) {
    if (cell.y == 0) return;
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of this stage are done
        mbar_expect_tx(st.bar[s], cell.y * 48u);
        tma_load_1d(st.buf[s], G.recs + 3 * (size_t)cell.x, cell.y * 48u, st.bar[s]);
    }
}

// all lanes: wait for stage s, test its records (lane k -> record k, k+32), reduce, update (t, hit)
template <bool FMA>
PT_DEV void coop_test(CoopStage &st, int s, uint2 cell, V3 o, V3 d, float &t, int &hit, int lane, Counters &cnt) {
    if (cell.y == 0) return;
    mbar_wait(st.bar[s], st.parity[s]);
    st.parity[s] ^= 1u;
    float best = t;
    int bestk = 0x7fffffff;
    for (uint32_t k = lane; k < cell.y; k += 32) {
        const float4 *rec = st.buf[s] + 3 * k;
        float r = t;
        if (tri_test<FMA>(rec[0], rec[1], rec[2], o, d, r)) {
            // within one lane k increases, so "<" keeps the first of equal distances
            if (r < best) { best = r; bestk = (int)k; }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int ok = __shfl_xor_sync(0xffffffffu, bestk, off);
        if (ok != 0x7fffffff && (bestk == 0x7fffffff || ob < best || (ob == best && ok < bestk))) { best = ob; bestk = ok; }
    }
    if (bestk != 0x7fffffff) { t = best; hit = hit_make(HIT_TRI, (int)cell.x + bestk); }
    if (lane == 0) { cnt.cells++; cnt.gtri += cell.y; cnt.btests += cell.y; }
    __syncwarp();
}

template <bool FMA>
PT_DEV void trace_grid_coop(const GridDev &G, CoopStage &st, V3 o, V3 d, float &t, int &hit, int lane, Counters &cnt) {
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
    if (t0 > t1) return;
    bool inside = o.x >= G.bmin[0] && o.x <= G.bmax[0] && o.y >= G.bmin[1] && o.y <= G.bmax[1] &&
                  o.z >= G.bmin[2] && o.z <= G.bmax[2];
    float next[3], dl[3];
    int idx[3], step[3], stop[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float p = inside ? oo[a] : A::madd(dd[a], t0, oo[a]);
        int hi = G.res[a] - 1;
        int v = f2i_rz_sat(A::div(A::sub(p, G.bmin[a]), G.cell[a]));
        idx[a] = min(max(v, 0), hi);
        dl[a] = A::div(A::sub(tX[a], tE[a]), __int2float_rn(G.res[a]));
        bool pos = dd[a] > 0.0f;
        next[a] = A::madd(__int2float_rn(pos ? idx[a] + 1 : G.res[a] - idx[a]), dl[a], tE[a]);
        step[a] = pos ? 1 : -1;
        stop[a] = pos ? G.res[a] : -1;
    }
    const int rx = G.res[0], rxy = G.res[0] * G.res[1];
    int s = 0;
    uint2 cell = __ldg(&G.cells[(size_t)idx[2] * rxy + (size_t)idx[1] * rx + idx[0]]);
    coop_issue(G, st, s, cell, lane);
    for (;;) {
        // advance the DDA one step (axis choice is independent of t) and prefetch that cell
        int kk = ((next[0] < next[1]) << 2) + ((next[0] < next[2]) << 1) + (next[1] < next[2]);
        int axis = (0x00221212u >> (4 * kk)) & 0xF;
        float lim;
        bool at_end;
        if (axis == 0)      { next[0] = A::add(next[0], dl[0]); lim = next[0]; idx[0] += step[0]; at_end = idx[0] == stop[0]; }
        else if (axis == 1) { next[1] = A::add(next[1], dl[1]); lim = next[1]; idx[1] += step[1]; at_end = idx[1] == stop[1]; }
        else                { next[2] = A::add(next[2], dl[2]); lim = next[2]; idx[2] += step[2]; at_end = idx[2] == stop[2]; }
        uint2 ncell = make_uint2(0u, 0u);
        if (!at_end) {
            ncell = __ldg(&G.cells[(size_t)idx[2] * rxy + (size_t)idx[1] * rx + idx[0]]);
            coop_issue(G, st, s ^ 1, ncell, lane);
        }
        coop_test<FMA>(st, s, cell, o, d, t, hit, lane, cnt);
        if (cell.y == 0 && lane == 0) cnt.cells++;
        const bool stop_now = (t < lim) || at_end;      // grid:194-197 (t compared AFTER the increment)
        if (stop_now) {
            if (!at_end && ncell.y) {                    // drain the speculative prefetch before the stage is reused
                mbar_wait(st.bar[s ^ 1], st.parity[s ^ 1]);
                st.parity[s ^ 1] ^= 1u;
            }
            break;
        }
        cell = ncell;
        s ^= 1;
    }
}

template <bool FMA>
PT_DEV int trace_ray_coop(const LaunchArgs &P, const SceneBlock *S, CoopStage &st, V3 o, V3 d, float &t, int lane, Counters &cnt) {
    if (lane == 0) cnt.rays++;
    int hit = HIT_NONE;
    trace_analytic<FMA, true>(P.ap, S, o, d, t, hit);       // warp-uniform (every lane computes the same)
    trace_grid_coop<FMA>(P.grid, st, o, d, t, hit, lane, cnt);
    return hit;
}

template <bool FMA>
__global__ void __launch_bounds__(128) k_grid_tma(const __grid_constant__ LaunchArgs P, uint32_t nitems, uint32_t *work_counter) {
    typedef Ar<FMA> A;
    __shared__ __align__(128) unsigned char s_buf[4][2][PT_TMA_STAGE_BYTES + 96];
    __shared__ __align__(8) uint64_t s_bar[4][2];
    const SceneBlock *S = &c_scene;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    CoopStage st;
    for (int s = 0; s < 2; ++s) {
        st.buf[s] = reinterpret_cast<float4 *>(&s_buf[warp][s][0]);
        st.bar[s] = &s_bar[warp][s];
        st.parity[s] = 0;
    }
    if (lane == 0) {
        mbar_init(st.bar[0], 1);
        mbar_init(st.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    Counters cnt = {0, 0, 0, 0, 0, 0};
    uint32_t w = blockIdx.x * 4 + warp;      // one work item = one pixel = one warp
    for (;;) {
        if (w >= nitems) break;
        int i, j;
        if (item_to_pixel(P, w, i, j)) {
            Rng rng = rng_seed(P.seeds, (uint32_t)(j * P.W + i));
            float cx = P.c0, cy = P.c0, cz = P.c0;
            for (int s = 0; s < P.spp; ++s) {
                V3 o, d;
                camera_ray<FMA>(P.cam, rng, i, j, o, d);
                if (lane == 0) cnt.samples++;
                float t = 1e9f;
                const int hit = trace_ray_coop<FMA>(P, S, st, o, d, t, lane, cnt);
                V3 c;
                if (hit == HIT_NONE) {
                    c = shade_sky<FMA>(d);
                } else {
                    const int m = hit_material(hit);
                    const V3 n = hit_normal<FMA, true>(P.ap, S, P.grid, hit, o, d, t);
                    const V3 X = A::vmadd(d, t, o);
                    float illum = 0.0f;
                    for (int l = 0; l < P.ap.nlights; ++l) {
                        float r0, r1;
                        rng_next(rng, r0, r1);
                        const float4 Lt = P.ap.lights[l];
                        V3 ld; float lam;
                        light_dir<FMA>(Lt, r0, r1, X, n, ld, lam);
                        if (lam < 0.0f) continue;
                        if (lane == 0) cnt.shadow++;
                        if (trace_ray_coop<FMA>(P, S, st, X, ld, t, lane, cnt) != HIT_NONE) continue;
                        illum = light_add<FMA>(Lt, X, lam, illum);
                    }
                    c = shade_material<FMA>(m, illum, X, n, d);
                }
                cx = A::madd(c.x, P.scale, cx); cy = A::madd(c.y, P.scale, cy); cz = A::madd(c.z, P.scale, cz);
            }
            if (lane == 0) {
                const size_t pix = (size_t)j * P.W + i;
                P.rgba[pix] = pack_rgba8_rz(cx, cy, cz, P.alpha);
                if (P.accum) P.accum[pix] = make_float4(cx, cy, cz, P.alpha);
                if (P.rng_out) P.rng_out[pix] = make_uint4(rng.x0, rng.x1, rng.c0, rng.c1);
            }
        }
        uint32_t nw = 0;
        if (lane == 0) nw = atomicAdd(work_counter, 1u);
        w = __shfl_sync(0xffffffffu, nw, 0);
    }
    __syncwarp();
    flush_counters(P, cnt, 0, P.ap.nsq + P.ap.nsp);
}

}  // namespace pt

int pt_launch_grid_tma(pt_ctx ctx, const pt_render_params *p, const pt::LaunchArgs &args) {
    using namespace pt;
    if (p->variant != PT_VARIANT_GRID) return pt_fail(1, "PT_KERNEL_GRID_TMA applies to the trianglegrid variant only");
    const bool fma = p->arith != PT_ARITH_SEPARATE;
    int rc = pt_bind_const_scene(ctx, fma ? PT_ARITH_FMA : PT_ARITH_SEPARATE);
    if (rc) return rc;
    const uint32_t tiles_x = (uint32_t)(args.W + 7) / 8, tiles_y = (uint32_t)(args.nrows + 3) / 4;
    const uint32_t nitems = tiles_x * tiles_y * 32u;
    uint32_t blocks = (uint32_t)ctx->sm_count * 8u;
    if (blocks * 4u > nitems) blocks = (nitems + 3) / 4;
    if (pt_ensure_scratch(ctx, 256)) return 1;
    uint32_t *counter = (uint32_t *)ctx->d_scratch;
    const uint32_t first_free = blocks * 4u;
    PT_CUDA(cudaMemcpyAsync(counter, &first_free, 4, cudaMemcpyHostToDevice, ctx->stream), "init work counter");
    if (fma) k_grid_tma<true><<<blocks, 128, 0, ctx->stream>>>(args, nitems, counter);
    else k_grid_tma<false><<<blocks, 128, 0, ctx->stream>>>(args, nitems, counter);
    PT_CUDA(cudaGetLastError(), "launch k_grid_tma");
    return 0;
}
