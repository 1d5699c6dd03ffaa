// pt_wavefront.cuh — WAVEFRONT pipeline (PT_KERNEL_WAVEFRONT): generate / intersect / shade / compact.
//
// One "pass" advances every pixel by one sample through separate kernels that communicate through
// SoA float4 buffers in HBM (coalesced 16-byte accesses, one element per thread):
//
//   wf_generate      rng -> primary ray                 rayO = (o.xyz, t=1e9)   rayD = (d.xyz, -)
//   wf_intersect     closest hit of a ray queue         rayO.w = t              rayD.w = hit code
//   wf_shade         miss: sky.  hit: X, n, material;   draws light jitter, pushes a shadow-ray entry into a
//                    COMPACTED queue (warp ballot + one atomicAdd per warp) or finishes the sample
//   wf_intersect     the shadow queue (origin X, direction sdir, bound t)
//   wf_shade_shadow  adds the light if unoccluded, moves to the next light (push) or finishes
//
// Lights are visited one queue generation at a time because the lmem family carries `t` from one shadow
// ray to the next (lmem:155,178).  Per-pixel RNG streams stay in order, so results are bit-identical to
// the megakernels.  NoDoF runs 64 passes (pass = local sample id li) into a 64-plane scratch image and a
// final kernel applies the reference's 8x8 reduction tree (nodof:253-274) — the reference's own structure.
// No host synchronisation inside a frame: shadow kernels are launched for the worst case and threads
// beyond the device-side queue length exit.
#pragma once
#include "pt_host.h"

namespace pt {

struct WfBuffers {
    uint32_t n;            // work-items per pass (pixels of the launch window)
    uint4 *rng;            // per pixel RNG state (pixel variants)
    float4 *color;         // per pixel accumulated colour
    float4 *rayO, *rayD;   // primary ray, later (t, hit code) in the .w lanes
    float4 *X;             // hit point, .w = accumulated illumination
    float4 *nrm;           // normal, .w = material aux (checker parity / facing ratio)
    float4 *sdir;          // shadow ray direction, .w = Lambert factor
    int2 *misc;            // .x = material, .y = light index in flight
    uint32_t *queue[2];    // compacted work-item ids
    uint32_t *qcount;      // queue lengths, one per generation
    float4 *tmp;           // NoDoF: 64 planes of n float4
};

PT_DEV bool wf_item(const LaunchArgs &P, uint32_t idx, int &i, int &j) {
    const int vr = (int)(idx / (uint32_t)P.W);
    i = (int)(idx - (uint32_t)vr * (uint32_t)P.W);
    j = map_row(P, vr);
    return vr < P.nrows && j < P.row_end;
}

// push `idx` into the output queue: one atomicAdd per warp (ballot + popc compaction)
PT_DEV void wf_push(bool push, uint32_t idx, uint32_t *queue, uint32_t *count) {
    const unsigned m = __ballot_sync(0xffffffffu, push);
    if (!m) return;
    const unsigned lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(count, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (push) queue[base + __popc(m & ((1u << lane) - 1u))] = idx;
}

template <bool FMA, bool NODOF>
__global__ void __launch_bounds__(256) wf_generate(const __grid_constant__ LaunchArgs P, WfBuffers B, int pass) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    int i, j;
    if (idx >= B.n || !wf_item(P, idx, i, j)) return;
    Rng rng;
    if (NODOF) {
        const int gi = 8 * i + (pass & 7), gj = 8 * j + (pass >> 3);
        rng = rng_seed(P.seeds, (uint32_t)(gj * (8 * P.W) + gi));
    } else if (pass == 0) {
        rng = rng_seed(P.seeds, (uint32_t)(j * P.W + i));
        B.color[idx] = make_float4(P.c0, P.c0, P.c0, P.alpha);
    } else {
        const uint4 s = B.rng[idx];
        rng.x0 = s.x; rng.x1 = s.y; rng.c0 = s.z; rng.c1 = s.w;
    }
    V3 o, d;
    camera_ray<FMA>(P.cam, rng, i, j, o, d);
    B.rayO[idx] = make_float4(o.x, o.y, o.z, 1e9f);
    B.rayD[idx] = make_float4(d.x, d.y, d.z, 0.0f);
    B.rng[idx] = make_uint4(rng.x0, rng.x1, rng.c0, rng.c1);
}

// SHADOW = false: all n primary rays.  SHADOW = true: the entries of `queue`.
template <bool FMA, bool CARRY, bool GRID, bool SHADOW>
__global__ void __launch_bounds__(256) wf_intersect(const __grid_constant__ LaunchArgs P, WfBuffers B, const uint32_t *queue,
                                                    const uint32_t *count) {
    const SceneBlock *S = &c_scene;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    Counters cnt = {0, 0, 0, 0, 0, 0};
    bool live = SHADOW ? tid < *count : tid < B.n;
    uint32_t idx = tid;
    if (live && SHADOW) idx = queue[tid];
    int i, j;
    if (live && !SHADOW) live = wf_item(P, idx, i, j);
    if (live) {
        const float4 ro = B.rayO[idx];
        V3 o, d;
        float t = ro.w;
        if (SHADOW) {
            const float4 x = B.X[idx], sd = B.sdir[idx];
            o = mk3(x.x, x.y, x.z); d = mk3(sd.x, sd.y, sd.z);
            cnt.shadow++;
        } else {
            const float4 rd = B.rayD[idx];
            o = mk3(ro.x, ro.y, ro.z); d = mk3(rd.x, rd.y, rd.z);
        }
        const int hit = trace_ray<FMA, CARRY, GRID>(P.ap, S, P.grid, o, d, t, cnt);
        B.rayO[idx].w = t;
        B.rayD[idx].w = __int_as_float(hit);
    }
    if (GRID) cnt.btests = cnt.gtri;        // trace_grid tests every record of the cells it visits
    flush_counters(P, cnt, GRID ? 0 : S->ntri_counted, P.ap.nsq + P.ap.nsp);
}

// Shared tail: visit lights from `l` on; either a shadow ray is needed (state stored, returns true) or the
// sample is finished with colour `c`.
template <bool FMA, bool CARRY>
PT_DEV bool wf_next_light(const LaunchArgs &P, WfBuffers &B, uint32_t idx, Rng &rng, V3 X, V3 n, int mat, float matf, float illum,
                          int l, V3 &c) {
    typedef Ar<FMA> A;
    while (l < P.ap.nlights) {
        float r0, r1;
        rng_next(rng, r0, r1);
        const float4 Lt = P.ap.lights[l];
        if (!CARRY && Lt.w == 0.0f) { ++l; continue; }
        V3 ld; float lam;
        light_dir<FMA>(Lt, r0, r1, X, n, ld, lam);
        if (lam < 0.0f) { ++l; continue; }
        B.X[idx] = make_float4(X.x, X.y, X.z, illum);
        B.sdir[idx] = make_float4(ld.x, ld.y, ld.z, lam);
        B.misc[idx] = make_int2(mat, l);
        return true;
    }
    if (illum > 1.0f) illum = 1.0f;
    illum = A::mul(illum, 0.25f);
    if (mat == 1) {
        const float i3 = A::mul(3.0f, illum);
        c = matf != 0.0f ? mk3(i3, illum, illum) : mk3(i3, i3, i3);
    } else if (mat == 3) {
        const float i2 = A::mul(2.0f, illum);
        c = mk3(i2, A::mul(3.0f, illum), i2);
    } else {
        c = mk3(matf, matf, matf);
    }
    return false;
}

template <bool FMA, bool NODOF>
PT_DEV void wf_finish(const LaunchArgs &P, WfBuffers &B, uint32_t idx, int pass, V3 c) {
    if (NODOF) {
        B.tmp[(size_t)pass * B.n + idx] = make_float4(__fmul_rn(c.x, 3.5f), __fmul_rn(c.y, 3.5f), __fmul_rn(c.z, 3.5f), 0.0f);
    } else {
        float4 col = B.color[idx];
        col.x = Ar<FMA>::madd(c.x, P.scale, col.x);
        col.y = Ar<FMA>::madd(c.y, P.scale, col.y);
        col.z = Ar<FMA>::madd(c.z, P.scale, col.z);
        B.color[idx] = col;
    }
}

// After the primary intersect.  FIRST = true: work-items are all pixels; false: entries of qin (shadow results).
template <bool FMA, bool CARRY, bool GRID, bool NODOF, bool FIRST>
__global__ void __launch_bounds__(256) wf_shade(const __grid_constant__ LaunchArgs P, WfBuffers B, int pass, const uint32_t *qin,
                                                const uint32_t *qin_count, uint32_t *qout, uint32_t *qout_count) {
    typedef Ar<FMA> A;
    const SceneBlock *S = &c_scene;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    Counters cnt = {0, 0, 0, 0, 0, 0};
    bool live = FIRST ? tid < B.n : tid < *qin_count;
    uint32_t idx = tid;
    if (live && !FIRST) idx = qin[tid];
    int i, j;
    if (live && FIRST) live = wf_item(P, idx, i, j);
    bool push = false;
    if (live) {
        const uint4 s = B.rng[idx];
        Rng rng; rng.x0 = s.x; rng.x1 = s.y; rng.c0 = s.z; rng.c1 = s.w;
        const float4 ro = B.rayO[idx], rd = B.rayD[idx];
        const int hit = __float_as_int(rd.w);
        V3 c = mk3(0.f, 0.f, 0.f);
        if (FIRST) {
            cnt.samples++;
            const V3 o = mk3(ro.x, ro.y, ro.z), d = mk3(rd.x, rd.y, rd.z);
            if (hit == HIT_NONE) {
                c = shade_sky<FMA>(d);
            } else {
                const int mat = hit_material(hit);
                const V3 n = hit_normal<FMA, GRID>(P.ap, S, P.grid, hit, o, d, ro.w);
                const V3 X = A::vmadd(d, ro.w, o);
                float matf = 0.0f;
                if (mat == 1) {
                    const float yx = A::mul(X.x, 0.2f), yy = A::mul(X.y, 0.2f);
                    matf = (f2i_rz_sat(A::add(ceilf(yx), ceilf(yy))) & 1) ? 1.0f : 0.0f;
                } else if (mat == 4) {
                    const float fr = A::dot(n, mk3(-d.x, -d.y, -d.z));
                    matf = 0.0f < fr ? fr : 0.0f;
                }
                B.nrm[idx] = make_float4(n.x, n.y, n.z, matf);
                push = wf_next_light<FMA, CARRY>(P, B, idx, rng, X, n, mat, matf, 0.0f, 0, c);
            }
        } else {
            const float4 x = B.X[idx], nn = B.nrm[idx], sd = B.sdir[idx];
            const int2 ms = B.misc[idx];
            const V3 X = mk3(x.x, x.y, x.z);
            float illum = x.w;
            if (hit == HIT_NONE) illum = light_add<FMA>(P.ap.lights[ms.y], X, sd.w, illum);
            push = wf_next_light<FMA, CARRY>(P, B, idx, rng, X, mk3(nn.x, nn.y, nn.z), ms.x, nn.w, illum, ms.y + 1, c);
        }
        B.rng[idx] = make_uint4(rng.x0, rng.x1, rng.c0, rng.c1);
        if (!push) wf_finish<FMA, NODOF>(P, B, idx, pass, c);
    }
    wf_push(push, idx, qout, qout_count);
    flush_counters(P, cnt, 0, 0);
}

template <bool NODOF>
__global__ void __launch_bounds__(256) wf_resolve(const __grid_constant__ LaunchArgs P, WfBuffers B) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    int i, j;
    if (idx >= B.n || !wf_item(P, idx, i, j)) return;
    const size_t pix = (size_t)j * P.W + i;
    float x, y, z;
    if (NODOF) {
        float vx[64], vy[64], vz[64];
#pragma unroll
        for (int li = 0; li < 64; ++li) {
            const float4 v = B.tmp[(size_t)li * B.n + idx];
            vx[li] = v.x; vy[li] = v.y; vz[li] = v.z;
        }
#pragma unroll
        for (int working = 32; working > 0; working >>= 1)
#pragma unroll
            for (int li = 0; li < working; ++li) {
                vx[li] = __fadd_rn(vx[li], vx[li + working]);
                vy[li] = __fadd_rn(vy[li], vy[li + working]);
                vz[li] = __fadd_rn(vz[li], vz[li + working]);
            }
        x = __fadd_rn(vx[0], 13.0f); y = __fadd_rn(vy[0], 13.0f); z = __fadd_rn(vz[0], 13.0f);
    } else {
        const float4 c = B.color[idx];
        x = c.x; y = c.y; z = c.z;
        if (P.rng_out) P.rng_out[pix] = B.rng[idx];
    }
    P.rgba[pix] = pack_rgba8_rz(x, y, z, NODOF ? 255.0f : P.alpha);
    if (P.accum) P.accum[pix] = make_float4(x, y, z, NODOF ? 255.0f : P.alpha);
}

// NoDoF keeps no per-pixel stream; its per-sample final RNG states are written by this helper pass.
__global__ void wf_store_sample_rng(const __grid_constant__ LaunchArgs P, WfBuffers B, int pass) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    int i, j;
    if (idx >= B.n || !wf_item(P, idx, i, j) || !P.rng_out) return;
    const int gi = 8 * i + (pass & 7), gj = 8 * j + (pass >> 3);
    P.rng_out[(size_t)gj * (8 * P.W) + gi] = B.rng[idx];
}

template <int VARIANT, bool FMA>
static int launch_wavefront_v(pt_ctx ctx, const LaunchArgs &args) {
    constexpr bool CARRY = VARIANT != PT_VARIANT_BASE;
    constexpr bool GRID = VARIANT == PT_VARIANT_GRID;
    constexpr bool NODOF = VARIANT == PT_VARIANT_NODOF;
    const uint32_t n = (uint32_t)args.W * (uint32_t)args.nrows;
    const int passes = NODOF ? 64 : args.spp;
    const int nl = args.ap.nlights;
    // carve the scratch arena
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_rng = carve((size_t)n * 16), o_col = carve((size_t)n * 16), o_ro = carve((size_t)n * 16), o_rd = carve((size_t)n * 16);
    const size_t o_x = carve((size_t)n * 16), o_n = carve((size_t)n * 16), o_sd = carve((size_t)n * 16), o_ms = carve((size_t)n * 8);
    const size_t o_q0 = carve((size_t)n * 4), o_q1 = carve((size_t)n * 4), o_qc = carve(64 * 4);
    const size_t o_tmp = NODOF ? carve((size_t)n * 16 * 64) : 0;
    if (pt_ensure_scratch(ctx, off)) return 1;
    char *base = (char *)ctx->d_scratch;
    WfBuffers B;
    B.n = n;
    B.rng = (uint4 *)(base + o_rng); B.color = (float4 *)(base + o_col); B.rayO = (float4 *)(base + o_ro); B.rayD = (float4 *)(base + o_rd);
    B.X = (float4 *)(base + o_x); B.nrm = (float4 *)(base + o_n); B.sdir = (float4 *)(base + o_sd); B.misc = (int2 *)(base + o_ms);
    B.queue[0] = (uint32_t *)(base + o_q0); B.queue[1] = (uint32_t *)(base + o_q1); B.qcount = (uint32_t *)(base + o_qc);
    B.tmp = NODOF ? (float4 *)(base + o_tmp) : nullptr;
    const unsigned blocks = (n + 255) / 256;
    cudaStream_t st = ctx->stream;
    // A frame is ~8 small launches per pass x 64+ passes and is launch-bound.  None of them needs the host (queue
    // lengths stay on the device), so the whole frame is captured ONCE into a CUDA graph and replayed with a single
    // cudaGraphLaunch for as long as the launch arguments, variant, policy and scratch arena are unchanged
    // (PT_WF_GRAPH=0 keeps the plain launches).
    static int use_graph = -1;
    if (use_graph < 0) { const char *e = getenv("PT_WF_GRAPH"); use_graph = e ? atoi(e) : 1; }
    if (use_graph && ctx->wf_exec && ctx->wf_key_args && ctx->wf_key_variant == VARIANT && ctx->wf_key_fma == (int)FMA &&
        ctx->wf_key_scratch == ctx->d_scratch && memcmp(ctx->wf_key_args, &args, sizeof(LaunchArgs)) == 0) {
        PT_CUDA(cudaGraphLaunch(ctx->wf_exec, st), "replay wavefront graph");
        return 0;
    }
    bool capturing = use_graph != 0;
    if (capturing && cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        (void)cudaGetLastError();          // e.g. the legacy default stream cannot be captured: plain launches instead
        capturing = false;
    }
    for (int pass = 0; pass < passes; ++pass) {
        cudaMemsetAsync(B.qcount, 0, 64 * 4, st);
        wf_generate<FMA, NODOF><<<blocks, 256, 0, st>>>(args, B, pass);
        wf_intersect<FMA, CARRY, GRID, false><<<blocks, 256, 0, st>>>(args, B, nullptr, nullptr);
        wf_shade<FMA, CARRY, GRID, NODOF, true><<<blocks, 256, 0, st>>>(args, B, pass, nullptr, nullptr, B.queue[0], B.qcount + 0);
        for (int g = 0; g < nl; ++g) {
            uint32_t *qin = B.queue[g & 1], *qout = B.queue[(g + 1) & 1];
            wf_intersect<FMA, CARRY, GRID, true><<<blocks, 256, 0, st>>>(args, B, qin, B.qcount + g);
            wf_shade<FMA, CARRY, GRID, NODOF, false><<<blocks, 256, 0, st>>>(args, B, pass, qin, B.qcount + g, qout, B.qcount + g + 1);
        }
        if (NODOF && args.rng_out) wf_store_sample_rng<<<blocks, 256, 0, st>>>(args, B, pass);
    }
    wf_resolve<NODOF><<<blocks, 256, 0, st>>>(args, B);
    if (!capturing) {
        PT_CUDA(cudaGetLastError(), "launch wavefront");
        return 0;
    }
    cudaGraph_t graph = nullptr;
    PT_CUDA(cudaStreamEndCapture(st, &graph), "end wavefront capture");
    if (ctx->wf_exec) { cudaGraphExecDestroy(ctx->wf_exec); ctx->wf_exec = nullptr; }
    cudaError_t ie = cudaGraphInstantiate(&ctx->wf_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { ctx->wf_exec = nullptr; return pt_cuda_fail(ie, "instantiate wavefront graph"); }
    if (!ctx->wf_key_args) ctx->wf_key_args = (LaunchArgs *)malloc(sizeof(LaunchArgs));
    memcpy(ctx->wf_key_args, &args, sizeof(LaunchArgs));
    ctx->wf_key_variant = VARIANT; ctx->wf_key_fma = (int)FMA; ctx->wf_key_scratch = ctx->d_scratch;
    PT_CUDA(cudaGraphLaunch(ctx->wf_exec, st), "launch wavefront graph");
    return 0;
}

template <bool FMA>
static int launch_wavefront_a(pt_ctx ctx, int variant, const LaunchArgs &args) {
    switch (variant) {
        case PT_VARIANT_BASE: return launch_wavefront_v<PT_VARIANT_BASE, FMA>(ctx, args);
        case PT_VARIANT_LMEM: return launch_wavefront_v<PT_VARIANT_LMEM, FMA>(ctx, args);
        case PT_VARIANT_GRID: return launch_wavefront_v<PT_VARIANT_GRID, FMA>(ctx, args);
        case PT_VARIANT_NODOF: return launch_wavefront_v<PT_VARIANT_NODOF, FMA>(ctx, args);
    }
    return pt_fail(1, "unknown variant %d", variant);
}

}  // namespace pt

int pt_launch_wavefront(pt_ctx ctx, const pt_render_params *p, const pt::LaunchArgs &args) {
    using namespace pt;
    const bool fma = p->arith != PT_ARITH_SEPARATE;
    int rc = pt_bind_const_scene(ctx, fma ? PT_ARITH_FMA : PT_ARITH_SEPARATE);   // wavefront kernels read the __constant__ scene
    if (rc) return rc;
    return fma ? launch_wavefront_a<true>(ctx, p->variant, args) : launch_wavefront_a<false>(ctx, p->variant, args);
}
