// pt_gridstream.cuh — trianglegrid variant, PT_KERNEL_GRID_STREAM: ray REGENERATION AT CELL GRANULARITY.
//
// Why.  In the megakernel a warp walks 32 DDA traversals in lock step and every lane waits for the longest one:
// on the 1 M-triangle soup the traversal loop runs with 13.5 of 32 lanes (ncu source view, profiles/r1_12), the
// triangle tests with 10.  Regenerating whole rays (k_sm_pixel) does not help — a "step" is still a complete
// traversal.  Here the unit of lock-step work is ONE CELL VISIT:
//   * every lane carries its ray, its DDA cursor and its pixel/sample state machine (as k_sm_pixel's Lane);
//   * the warp alternates between a TRAVERSAL phase — all lanes with a live traversal visit one cell per
//     iteration — and a REFILL phase that is entered as soon as fewer than STREAM_THRESH lanes are still
//     traversing: lanes whose ray ended shade it, start their next ray (shadow ray, next sample, or a new pixel
//     fetched with one warp-aggregated atomicAdd), run the analytic tests and the grid entry, and join the
//     traversal again;
//   * the DDA keeps the megakernel's software pipelining: the word of the NEXT cell is requested one visit ahead.
// Per ray the operations and their order are those of trace_grid / Sample, so results stay bit-identical
// (grid:102-201, 203-283, 348-381); only which rays share a warp at a given moment changes.
//
// MEASURED (B200, 1 M-triangle soup, 1920x1080, 4 spp): 10.3 ms against 4.57 ms for the megakernel and 8.75 ms for
// ray-granular regeneration (k_sm_pixel); refill thresholds 4..28 and batches 1..32 all land at 10.8-13 ms.  Keeping
// every lane busy does not pay here: the lanes of a megakernel warp are an 8x4 pixel tile in the same phase, so
// their rays walk (nearly) the same cells and the cell / record loads coalesce into broadcasts, while regenerated
// lanes hold unrelated rays (memory divergence), and the cursor state costs 86 registers (20 warps/SM instead of
// 32) in a latency-bound loop.  Kept as a selectable flavour and as evidence; PT_KERNEL_AUTO never picks it.
#pragma once
#include "pt_persistent.cuh"

namespace pt {

#define STREAM_THRESH 20   // refill when fewer lanes than this are traversing
#define STREAM_BATCH 6     // cell visits between two looks at the refill condition

struct Cursor {            // DDA state of one ray (grid:157-198)
    float nx, ny, nz;      // next[]: parametric distance of the next cell boundary per axis
    float dx, dy, dz;      // delta[]
    int ix, iy, iz;        // cell index
    uint2 cell, ncell;     // (first record, count) of the current cell and of the one the pending step leads to
    float lim;             // next[axis] after the pending step: the bound *t is compared with when leaving `cell`
    bool at_end;           // the pending step leaves the grid
};

// the DDA step that follows the current cell: axis choice, next/idx update, request of the following cell's word
template <bool FMA>
PT_DEV void cursor_step(const GridDev &G, V3 d, Cursor &C) {
    typedef Ar<FMA> A;
    const int kk = ((C.nx < C.ny) << 2) + ((C.nx < C.nz) << 1) + (C.ny < C.nz);
    const int axis = (0x00221212u >> (4 * kk)) & 0xF;          // the reference's LUT {2,1,2,1,2,2,0,0}
    if (axis == 0)      { C.nx = A::add(C.nx, C.dx); C.lim = C.nx; C.ix += d.x > 0.0f ? 1 : -1; C.at_end = C.ix == (d.x > 0.0f ? G.res[0] : -1); }
    else if (axis == 1) { C.ny = A::add(C.ny, C.dy); C.lim = C.ny; C.iy += d.y > 0.0f ? 1 : -1; C.at_end = C.iy == (d.y > 0.0f ? G.res[1] : -1); }
    else                { C.nz = A::add(C.nz, C.dz); C.lim = C.nz; C.iz += d.z > 0.0f ? 1 : -1; C.at_end = C.iz == (d.z > 0.0f ? G.res[2] : -1); }
    C.ncell = make_uint2(0u, 0u);
    if (!C.at_end) C.ncell = __ldg(&G.cells[((size_t)C.iz * G.res[1] + (size_t)C.iy) * G.res[0] + C.ix]);
}

// slab test + DDA initialisation (first half of trace_grid); false: the ray misses the box
template <bool FMA>
PT_DEV bool cursor_enter(const GridDev &G, V3 o, V3 d, Cursor &C) {
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
    C.nx = next[0]; C.ny = next[1]; C.nz = next[2];
    C.dx = dl[0]; C.dy = dl[1]; C.dz = dl[2];
    C.ix = idx[0]; C.iy = idx[1]; C.iz = idx[2];
    C.cell = __ldg(&G.cells[((size_t)C.iz * G.res[1] + (size_t)C.iy) * G.res[0] + C.ix]);
    cursor_step<FMA>(G, d, C);
    return true;
}

// One cell visit: the cell's triangles in order, then the reference's termination test; true: keep traversing.
template <bool FMA>
PT_DEV bool cursor_visit(const GridDev &G, V3 o, V3 d, float &t, int &hit, Cursor &C, Counters &cnt) {
    cnt.cells++;
    cnt.gtri += C.cell.y;
    cnt.btests += C.cell.y;
    const float4 *rec = G.recs + 3 * (size_t)C.cell.x;
    for (uint32_t k = 0; k < C.cell.y; ++k, rec += 3) {
        float4 ra = __ldg(rec), rb = __ldg(rec + 1), rc = __ldg(rec + 2);
        if (tri_test<FMA>(ra, rb, rc, o, d, t)) hit = hit_make(HIT_TRI, (int)(C.cell.x + k));
    }
    if (t < C.lim || C.at_end) return false;                   // t compared AFTER the increment (grid:194-195)
    C.cell = C.ncell;
    cursor_step<FMA>(G, d, C);
    return true;
}

template <bool FMA, int MEM>
__global__ void __launch_bounds__(128, 5) k_stream_grid(const __grid_constant__ LaunchArgs P, uint32_t nitems, uint32_t *work_counter) {
    typedef Ar<FMA> A;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SceneBlock *S = (MEM == PT_SCENE_SMEM) ? stage_scene_smem(P, smem_raw) : &c_scene;
    const unsigned lane = threadIdx.x & 31;
    Counters cnt = {0, 0, 0, 0, 0, 0};
    Lane L;
    L.phase = 0; L.l = 0; L.mat = 0; L.illum = 0.f; L.lam = 0.f; L.matf = 0.f; L.t = 1e9f;
    L.o = L.d = L.X = L.n = mk3(0.f, 0.f, 0.f);
    L.px = L.py = 0;
    L.rng = rng_seed(P.seeds, 0u);
    Cursor C;
    C.nx = C.ny = C.nz = C.dx = C.dy = C.dz = C.lim = 0.f;
    C.ix = C.iy = C.iz = 0;
    C.cell = C.ncell = make_uint2(0u, 0u);
    C.at_end = true;
    float cx = P.c0, cy = P.c0, cz = P.c0;
    int s = 0, hit = HIT_NONE;
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    bool have = false;      // lane owns a pixel
    bool want = true;       // lane needs a (new) work item ...
    bool fetch = false;     // ... and must first draw a fresh index from the global counter
    bool trav = false;      // a traversal is in flight
    bool done = false;      // a ray has ended and its result (hit, L.t) waits to be consumed
    const int thresh = P.scatter_mul ? (int)(P.scatter_mul & 0xff) : STREAM_THRESH, batch = P.scatter_mul ? (int)(P.scatter_mul >> 8) : STREAM_BATCH;
    for (;;) {
        // ------------------------------------------------------------------------------------ REFILL
        if (__popc(__ballot_sync(0xffffffffu, trav)) < thresh) {
            for (int rep = 0; rep < 4; ++rep) {
                // work items for lanes without a pixel (as k_sm_pixel)
                while (__any_sync(0xffffffffu, want)) {
                    const unsigned need = __ballot_sync(0xffffffffu, want && fetch);
                    if (need) {
                        const int leader = __ffs(need) - 1;
                        uint32_t base = 0;
                        if ((int)lane == leader) base = atomicAdd(work_counter, (uint32_t)__popc(need));
                        base = __shfl_sync(0xffffffffu, base, leader);
                        if (want && fetch) w = base + __popc(need & ((1u << lane) - 1u));
                        fetch = false;
                    }
                    if (want) {
                        int i, j;
                        if (w >= nitems) {
                            want = false;                      // queue exhausted
                        } else if (item_to_pixel(P, w, i, j)) {
                            L.px = i; L.py = j;
                            L.rng = rng_seed(P.seeds, (uint32_t)(j * P.W + i));
                            L.phase = 0; s = 0; cx = cy = cz = P.c0;
                            have = true; want = false; done = false;
                        } else {
                            fetch = true;                      // item lies outside the image: draw another
                        }
                    }
                }
                const bool act = have && !trav;
                if (!__any_sync(0xffffffffu, act)) break;
                if (act) {
                    bool start = true;                         // start a ray at the end of this step
                    if (done) {                                // ---- consume the result of the finished ray (Sample, grid:203-283)
                        done = false;
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
                            if (++s == P.spp) {
                                const size_t pix = (size_t)L.py * P.W + L.px;
                                P.rgba[pix] = pack_rgba8_rz(cx, cy, cz, P.alpha);
                                if (P.accum) P.accum[pix] = make_float4(cx, cy, cz, P.alpha);
                                if (P.rng_out) P.rng_out[pix] = make_uint4(L.rng.x0, L.rng.x1, L.rng.c0, L.rng.c1);
                                have = false; want = true; fetch = true;
                                start = false;
                            }
                        }
                    }
                    if (start) {                               // ---- begin the next ray (TraceRay up to the grid entry, grid:102-176)
                        if (L.phase == 0) {
                            camera_ray<FMA>(P.cam, L.rng, L.px, L.py, L.o, L.d);
                            L.t = 1e9f;                        // grid:222
                        }
                        cnt.rays++;
                        hit = HIT_NONE;
                        trace_analytic<FMA, true>(P.ap, S, L.o, L.d, L.t, hit);
                        trav = cursor_enter<FMA>(P.grid, L.o, L.d, C);
                        done = !trav;
                    }
                }
            }
        }
        if (!__any_sync(0xffffffffu, trav || have || want)) break;
        // ------------------------------------------------------------------------------------ TRAVERSAL
        for (int it = 0; it < batch; ++it) {
            if (trav) {
                trav = cursor_visit<FMA>(P.grid, L.o, L.d, L.t, hit, C, cnt);
                done = !trav;
            }
            if (__popc(__ballot_sync(0xffffffffu, trav)) < thresh) break;
        }
    }
    flush_counters(P, cnt, 0, P.ap.nsq + P.ap.nsp);
}

template <bool FMA, int MEM>
static int launch_stream_grid(pt_ctx ctx, const LaunchArgs &args_in) {
    LaunchArgs args = args_in;
    args.scatter_mul = 0;
    if (getenv("PT_STREAM")) args.scatter_mul = (uint32_t)atoi(getenv("PT_STREAM"));   // tuning sweep: thresh | batch << 8
    const uint32_t tiles_x = (uint32_t)(args.W + 7) / 8, tiles_y = (uint32_t)(args.nrows + 3) / 4;
    const uint32_t nitems = tiles_x * tiles_y * 32u;
    const size_t smem = MEM == PT_SCENE_SMEM ? (size_t)args.scene_bytes : 0;
    auto kern = k_stream_grid<FMA, MEM>;
    int per_sm = 0;
    PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem), "occupancy query");
    if (per_sm < 1) per_sm = 1;
    uint32_t blocks = (uint32_t)(ctx->sm_count * per_sm);
    const uint32_t need_blocks = (nitems + 127) / 128;
    if (blocks > need_blocks || getenv("PT_STREAM_ALL")) blocks = need_blocks;   // PT_STREAM_ALL: one item per lane, no regeneration (each lane keeps its pixel)
    if (pt_ensure_scratch(ctx, 256)) return 1;
    uint32_t *counter = (uint32_t *)ctx->d_scratch;
    const uint32_t first_free = blocks * 128u;       // items [0, first_free) are the initial assignment
    PT_CUDA(cudaMemcpyAsync(counter, &first_free, 4, cudaMemcpyHostToDevice, ctx->stream), "init work counter");
    kern<<<blocks, 128, smem, ctx->stream>>>(args, nitems, counter);
    PT_CUDA(cudaGetLastError(), "launch k_stream_grid");
    return 0;
}

}  // namespace pt

int pt_launch_stream_grid(pt_ctx ctx, const pt_render_params *p, const pt::LaunchArgs &args) {
    using namespace pt;
    const bool fma = p->arith != PT_ARITH_SEPARATE;
    if (p->scene_mem == PT_SCENE_SMEM)
        return fma ? launch_stream_grid<true, PT_SCENE_SMEM>(ctx, args) : launch_stream_grid<false, PT_SCENE_SMEM>(ctx, args);
    int rc = pt_bind_const_scene(ctx, fma ? PT_ARITH_FMA : PT_ARITH_SEPARATE);
    if (rc) return rc;
    return fma ? launch_stream_grid<true, PT_SCENE_CONST>(ctx, args) : launch_stream_grid<false, PT_SCENE_CONST>(ctx, args);
}
