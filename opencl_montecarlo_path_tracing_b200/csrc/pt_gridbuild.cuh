// pt_gridbuild.cuh — deterministic uniform-grid build on the device.
//
// Replaces kernel initTrianglesGrid (grid:311-330) + atomic_addTriangle (grid:285-289).  The reference
// bins with atomic_inc, so the order of a cell's entries changes from run to run and `nels` keeps
// counting past the 62-entry capacity (CellIntersect then reads out of bounds).  Here:
//   1. count  : one thread per triangle adds 1 to every overlapped cell (same cell range arithmetic),
//   2. scan   : exclusive prefix sum of the raw counts,
//   3. fill   : triangle ids scattered into their cell's raw segment (arbitrary order),
//   4. sort   : every cell sorts its segment ascending => triangle-id order, i.e. exactly what the
//               reference's serial initTrianglesGrid_host (..._trianglegrid/CLSuperPathTracer.c:233-265)
//               produces; the first `cap` (62) entries are kept,
//   5. scan + emit : capped CSR (cell -> first record, count) and CONTIGUOUS per-cell triangle records
//               (e2, e0, v0 as 3 x float4), so a traversal step reads one dense block instead of
//               chasing 16-bit indices into a triangle array.  32-bit ids: no 65536-triangle limit.
#pragma once
#include <float.h>

#include "pt_host.h"

namespace pt {

PT_DEV void tri_cell_range(const float *t, const GridDev &G, int lo[3], int hi[3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float mn = cl_fmin(t[a], cl_fmin(t[4 + a], t[8 + a]));
        float mx = cl_fmax(t[a], cl_fmax(t[4 + a], t[8 + a]));
        int l = f2i_rz_sat(__fdiv_rn(__fsub_rn(mn, G.bmin[a]), G.cell[a]));
        int h = f2i_rz_sat(__fdiv_rn(__fsub_rn(mx, G.bmin[a]), G.cell[a]));
        lo[a] = min(max(l, 0), G.res[a] - 1);
        hi[a] = min(max(h, 0), G.res[a] - 1);
    }
}

// Small sphere around a triangle: the circumscribed circle's centre, or — obtuse triangles — the midpoint of the longest
// edge, in plain float (any centre is valid, only the tightness depends on it); the radius is the largest vertex
// distance from that centre, rounded up and inflated by 1 % + 0.01 exactly like the brute-force mesh / cluster spheres
// (ptcuda.cu pt_set_scene), which dwarfs the float rounding of the distances.  Degenerate or non-finite input ->
// radius +inf: the filter then always passes and the exact test decides.  (FP64 would cost 30 ms per 1 M-triangle build.)
PT_DEV float4 tri_bound_sphere(const float *t) {
    const float ax = t[0], ay = t[1], az = t[2];
    const float ux = t[4] - ax, uy = t[5] - ay, uz = t[6] - az;      // B - A
    const float vx = t[8] - ax, vy = t[9] - ay, vz = t[10] - az;     // C - A
    const float uu = ux * ux + uy * uy + uz * uz, vv = vx * vx + vy * vy + vz * vz, uv = ux * vx + uy * vy + uz * vz;
    const float ww = uu + vv - 2.0f * uv;                             // |C - B|^2
    float cx, cy, cz;
    const float den = 2.0f * (uu * vv - uv * uv);                     // 2 |u x v|^2
    if (uu + vv <= ww)      { cx = 0.5f * (ux + vx); cy = 0.5f * (uy + vy); cz = 0.5f * (uz + vz); }   // angle at A >= 90: edge BC
    else if (uu + ww <= vv) { cx = 0.5f * vx; cy = 0.5f * vy; cz = 0.5f * vz; }                        // angle at B >= 90: edge AC
    else if (vv + ww <= uu) { cx = 0.5f * ux; cy = 0.5f * uy; cz = 0.5f * uz; }                        // angle at C >= 90: edge AB
    else if (den > 1e-12f * (uu * vv)) {
        const float s = vv * (uu - uv) / den, q = uu * (vv - uv) / den;   // circumcentre = A + s u + q v
        cx = s * ux + q * vx; cy = s * uy + q * vy; cz = s * uz + q * vz;
    } else                  { cx = (ux + vx) * (1.0f / 3.0f); cy = (uy + vy) * (1.0f / 3.0f); cz = (uz + vz) * (1.0f / 3.0f); }
    const float fx = ax + cx, fy = ay + cy, fz = az + cz;
    float r2 = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float dx = t[4 * k] - fx, dy = t[4 * k + 1] - fy, dz = t[4 * k + 2] - fz;
        r2 = fmaxf(r2, dx * dx + dy * dy + dz * dz);
    }
    const float rr = __fadd_ru(__fmul_ru(__fsqrt_ru(__fmul_ru(r2, 1.00001f)), 1.01f), 0.01f);
    if (!(rr < 1e30f) || !(fabsf(fx) < 1e30f) || !(fabsf(fy) < 1e30f) || !(fabsf(fz) < 1e30f))
        return make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));
    return make_float4(fx, fy, fz, rr);
}

__global__ void k_grid_count(const float *__restrict__ tris, int ntri, GridDev G, uint32_t *__restrict__ count, uint32_t *__restrict__ kmax_bits) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntri) return;
    {
        // max |e0||e2| over the mesh (rounded up), for the distance-proportional margin of the sphere filter; finite
        // non-negative floats order like their bit patterns, NaN / inf patterns order above every finite one
        const float *t = tris + 12 * (size_t)i;
        const float ux = t[4] - t[0], uy = t[5] - t[1], uz = t[6] - t[2], vx = t[8] - t[0], vy = t[9] - t[1], vz = t[10] - t[2];
        const float kk = __fmul_ru(__fsqrt_ru((ux * ux + uy * uy + uz * uz) * (vx * vx + vy * vy + vz * vz)), 1.00001f);
        uint32_t bits = __float_as_uint(kk) & 0x7fffffffu;
        if (bits > *(volatile uint32_t *)kmax_bits) atomicMax(kmax_bits, bits);
    }
    int lo[3], hi[3];
    tri_cell_range(tris + 12 * (size_t)i, G, lo, hi);
    for (int z = lo[2]; z <= hi[2]; ++z)
        for (int y = lo[1]; y <= hi[1]; ++y)
            for (int x = lo[0]; x <= hi[0]; ++x)
                atomicAdd(&count[(size_t)z * G.res[0] * G.res[1] + (size_t)y * G.res[0] + x], 1u);
}

__global__ void k_grid_fill(const float *__restrict__ tris, int ntri, GridDev G, const uint32_t *__restrict__ start,
                            uint32_t *__restrict__ cursor, uint32_t *__restrict__ refs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntri) return;
    int lo[3], hi[3];
    tri_cell_range(tris + 12 * (size_t)i, G, lo, hi);
    for (int z = lo[2]; z <= hi[2]; ++z)
        for (int y = lo[1]; y <= hi[1]; ++y)
            for (int x = lo[0]; x <= hi[0]; ++x) {
                size_t c = (size_t)z * G.res[0] * G.res[1] + (size_t)y * G.res[0] + x;
                uint32_t pos = atomicAdd(&cursor[c], 1u);
                refs[start[c] + pos] = (uint32_t)i;
            }
}

// Exclusive scan of n uint32 values, 3 passes (per-block scan, scan of block sums by one block, add).
// `out` receives n+1 offsets.  If cap > 0 the inputs are clamped to cap first.
constexpr int SCAN_BLOCK = 1024;

__global__ void k_scan_blocks(const uint32_t *__restrict__ in, size_t n, uint32_t cap, uint32_t *__restrict__ out,
                              uint32_t *__restrict__ block_sums) {
    __shared__ uint32_t warp_sums[32];
    size_t i = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    uint32_t v = i < n ? in[i] : 0u;
    if (cap && v > cap) v = cap;
    uint32_t x = v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += y;
        }
        warp_sums[lane] = w;
    }
    __syncthreads();
    uint32_t incl = x + (warp ? warp_sums[warp - 1] : 0u);
    if (i < n) out[i] = incl - v;
    if (threadIdx.x == SCAN_BLOCK - 1) block_sums[blockIdx.x] = incl;
}

__global__ void k_scan_sums(uint32_t *block_sums, int nblocks) {
    // single block, serial over chunks of 1024 (nblocks <= 2048 for 128^3 cells)
    __shared__ uint32_t carry;
    __shared__ uint32_t warp_sums[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nblocks; base += SCAN_BLOCK) {
        int i = base + threadIdx.x;
        uint32_t v = i < nblocks ? block_sums[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, w, off);
                if (lane >= off) w += y;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        uint32_t incl = x + (warp ? warp_sums[warp - 1] : 0u) + carry;
        if (i < nblocks) block_sums[i] = incl - v;   // exclusive
        __syncthreads();
        if (threadIdx.x == SCAN_BLOCK - 1) carry = incl;
        __syncthreads();
    }
}

__global__ void k_scan_add(uint32_t *__restrict__ out, size_t n, const uint32_t *__restrict__ block_sums,
                           const uint32_t *__restrict__ in, uint32_t cap) {
    size_t i = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    if (i < n) out[i] += block_sums[blockIdx.x];
    if (i == n - 1) {   // total goes to out[n]
        uint32_t v = in[i];
        if (cap && v > cap) v = cap;
        out[n] = out[i] + v;
    }
}

// One thread per cell: sort the raw segment ascending (insertion sort; segments are short), keep the
// first `cap` ids, emit capped refs + contiguous triangle records + the (first, count) cell word.
__global__ void k_grid_emit(const float *__restrict__ tris, size_t ncells, uint32_t cap,
                            const uint32_t *__restrict__ raw_start, uint32_t *__restrict__ raw_refs,
                            const uint32_t *__restrict__ cap_start, uint32_t *__restrict__ refs,
                            float4 *__restrict__ recs, float4 *__restrict__ sph, uint2 *__restrict__ cells) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    uint32_t b = raw_start[c], e = raw_start[c + 1];
    for (uint32_t i = b + 1; i < e; ++i) {
        uint32_t key = raw_refs[i];
        uint32_t j = i;
        while (j > b && raw_refs[j - 1] > key) { raw_refs[j] = raw_refs[j - 1]; --j; }
        raw_refs[j] = key;
    }
    uint32_t n = e - b;
    if (n > cap) n = cap;
    uint32_t first = cap_start[c];
    cells[c] = make_uint2(first, n);
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t id = raw_refs[b + k];
        refs[first + k] = id;
        const float *t = tris + 12 * (size_t)id;
        float4 *r = recs + 3 * (size_t)(first + k);
        float e0x = __fsub_rn(t[4], t[0]), e0y = __fsub_rn(t[5], t[1]), e0z = __fsub_rn(t[6], t[2]);
        float e2x = __fsub_rn(t[8], t[0]), e2y = __fsub_rn(t[9], t[1]), e2z = __fsub_rn(t[10], t[2]);
        r[0] = make_float4(e2x, e2y, e2z, e0x);
        r[1] = make_float4(e0y, e0z, t[0], t[1]);
        r[2] = make_float4(t[2], __uint_as_float(id), 0.f, 0.f);
        sph[first + k] = tri_bound_sphere(t);
    }
}

static int exclusive_scan(pt_ctx ctx, const uint32_t *in, size_t n, uint32_t cap, uint32_t *out, uint32_t *block_sums) {
    int nblocks = (int)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
    k_scan_blocks<<<nblocks, SCAN_BLOCK, 0, ctx->stream>>>(in, n, cap, out, block_sums);
    k_scan_sums<<<1, SCAN_BLOCK, 0, ctx->stream>>>(block_sums, nblocks);
    k_scan_add<<<nblocks, SCAN_BLOCK, 0, ctx->stream>>>(out, n, block_sums, in, cap);
    PT_CUDA(cudaGetLastError(), "grid scan");
    return 0;
}

// ---- VLP bounding box and VLP grid of CLSuperMetropolisPathTracer_vlpgrid ------------------------------------------------
// reduceMinAndMax_lmem / _nwg (metropolispathtracer.ocl:538-619) and initVLPsGrid (:621-647) are pure functions of a VLP
// buffer (x y z intensity); a virtual point light with intensity 0 is a dummy and is ignored, every other one reaches
// 16 sqrt(intensity) around its position.  Here: one reduction kernel (warp shuffles + one ordered atomic per block)
// instead of two work-group tree passes — min / max do not depend on the order; NaN boxes never win, as in the reference's
// isless / isgreater selects when they arrive from the partner lane — and the same deterministic count / scan / fill /
// sort / emit as the triangle grid, so a cell lists its lights in ascending index order (the reference appends with
// atomic_inc: any order; a serial run gives ascending).
PT_DEV bool vlp_reach(float4 v, float lo[3], float hi[3]) {
    if (v.w == 0.0f) return false;
    const float r = __fmul_rn(16.0f, __fsqrt_rn(v.w));
    lo[0] = __fsub_rn(v.x, r); lo[1] = __fsub_rn(v.y, r); lo[2] = __fsub_rn(v.z, r);
    hi[0] = __fadd_rn(v.x, r); hi[1] = __fadd_rn(v.y, r); hi[2] = __fadd_rn(v.z, r);
    return true;
}

// out[0..2] = keys of the minima, out[3..5] = keys of the maxima (ordered_key: monotone float -> uint)
__global__ void k_vlp_bounds(const float4 *__restrict__ vpl, int n, unsigned *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {FLT_MIN, FLT_MIN, FLT_MIN};     // the reference's dummy box
    if (i < n) {
        float l[3], h[3];
        if (vlp_reach(vpl[i], l, h))
            for (int a = 0; a < 3; ++a) { if (l[a] == l[a]) lo[a] = l[a]; if (h[a] == h[a]) hi[a] = h[a]; }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        unsigned kmin = ordered_key(lo[a]), kmax = ordered_key(hi[a]);
        kmin = __reduce_min_sync(0xffffffffu, kmin);
        kmax = __reduce_max_sync(0xffffffffu, kmax);
        if ((threadIdx.x & 31) == 0) { atomicMin(out + a, kmin); atomicMax(out + 3 + a, kmax); }
    }
}

PT_DEV bool vlp_cell_range(float4 v, const GridDev &G, int lo[3], int hi[3]) {
    float l[3], h[3];
    if (!vlp_reach(v, l, h)) return false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int cl = f2i_rz_sat(__fdiv_rn(__fsub_rn(l[a], G.bmin[a]), G.cell[a]));
        const int ch = f2i_rz_sat(__fdiv_rn(__fsub_rn(h[a], G.bmin[a]), G.cell[a]));
        lo[a] = min(max(cl, 0), G.res[a] - 1);
        hi[a] = min(max(ch, 0), G.res[a] - 1);
    }
    return true;
}

__global__ void k_vlp_count(const float4 *__restrict__ vpl, int n, GridDev G, uint32_t *__restrict__ count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo[3], hi[3];
    if (!vlp_cell_range(vpl[i], G, lo, hi)) return;
    for (int z = lo[2]; z <= hi[2]; ++z)
        for (int y = lo[1]; y <= hi[1]; ++y)
            for (int x = lo[0]; x <= hi[0]; ++x)
                atomicAdd(&count[(size_t)z * G.res[0] * G.res[1] + (size_t)y * G.res[0] + x], 1u);
}

__global__ void k_vlp_fill(const float4 *__restrict__ vpl, int n, GridDev G, const uint32_t *__restrict__ start,
                           uint32_t *__restrict__ cursor, uint32_t *__restrict__ refs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo[3], hi[3];
    if (!vlp_cell_range(vpl[i], G, lo, hi)) return;
    for (int z = lo[2]; z <= hi[2]; ++z)
        for (int y = lo[1]; y <= hi[1]; ++y)
            for (int x = lo[0]; x <= hi[0]; ++x) {
                size_t c = (size_t)z * G.res[0] * G.res[1] + (size_t)y * G.res[0] + x;
                uint32_t pos = atomicAdd(&cursor[c], 1u);
                refs[start[c] + pos] = (uint32_t)i;
            }
}

// one thread per cell: ascending light indices, the first `cap` kept
__global__ void k_vlp_emit(size_t ncells, uint32_t cap, const uint32_t *__restrict__ raw_start, uint32_t *__restrict__ raw_refs,
                           const uint32_t *__restrict__ cap_start, uint32_t *__restrict__ refs) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    uint32_t b = raw_start[c], e = raw_start[c + 1];
    for (uint32_t i = b + 1; i < e; ++i) {
        uint32_t key = raw_refs[i];
        uint32_t j = i;
        while (j > b && raw_refs[j - 1] > key) { raw_refs[j] = raw_refs[j - 1]; --j; }
        raw_refs[j] = key;
    }
    uint32_t n = e - b;
    if (n > cap) n = cap;
    const uint32_t first = cap_start[c];
    for (uint32_t k = 0; k < n; ++k) refs[first + k] = raw_refs[b + k];
}

static int grow_buf(void **ptr, size_t *capacity, size_t bytes, const char *what) {
    if (*ptr && *capacity >= bytes) return 0;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *capacity = 0;
    PT_CUDA(cudaMalloc(ptr, bytes), what);
    *capacity = bytes;
    return 0;
}

}  // namespace pt

// Bounding box of the context's VLP buffer (reduceMinAndMax_lmem + _nwg): vmin / vmax as the reference host reads them back
int pt_vlp_bounds_device(pt_ctx ctx, float vmin[4], float vmax[4]) {
    using namespace pt;
    if (grow_buf((void **)&ctx->d_vlp_keys, &ctx->vlp_keys_cap, 6 * sizeof(unsigned), "alloc VLP bounds")) return 1;
    unsigned init[6], keys[6];
    {
        const float fmx = FLT_MAX, fmn = FLT_MIN;
        unsigned u;
        memcpy(&u, &fmx, 4); init[0] = init[1] = init[2] = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        memcpy(&u, &fmn, 4); init[3] = init[4] = init[5] = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    }
    PT_CUDA(cudaMemcpyAsync(ctx->d_vlp_keys, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream), "init VLP bounds");
    if (ctx->nvpl > 0) {
        k_vlp_bounds<<<(ctx->nvpl + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_vpls, ctx->nvpl, ctx->d_vlp_keys);
        PT_CUDA(cudaGetLastError(), "VLP bounds");
    }
    PT_CUDA(cudaMemcpyAsync(keys, ctx->d_vlp_keys, sizeof(keys), cudaMemcpyDeviceToHost, ctx->stream), "read VLP bounds");
    PT_CUDA(cudaStreamSynchronize(ctx->stream), "sync VLP bounds");
    for (int a = 0; a < 6; ++a) {
        const unsigned k = keys[a], u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
        float f;
        memcpy(&f, &u, 4);
        (a < 3 ? vmin : vmax)[a % 3] = f;
    }
    vmin[3] = vmax[3] = 0.0f;
    return 0;
}

// initVLPsGrid on the context's VLP buffer: capped CSR (cell -> first ref, refs = light indices, ascending per cell)
int pt_vlp_grid_build_device(pt_ctx ctx, const pt_grid *g) {
    using namespace pt;
    const size_t ncells = (size_t)g->res[0] * g->res[1] * g->res[2];
    const uint32_t cap = g->max_refs_per_cell > 0 ? (uint32_t)g->max_refs_per_cell : 62u;
    const int n = ctx->nvpl;
    GridDev G;
    for (int a = 0; a < 3; ++a) { G.bmin[a] = g->box_min[a]; G.bmax[a] = g->box_max[a]; G.cell[a] = g->cell_size[a]; G.res[a] = g->res[a]; }
    G.cells = nullptr; G.cells_pad = nullptr; G.pad_sx = G.res[0] + 2; G.pad_sxy = G.pad_sx * (G.res[1] + 2); G.recs = nullptr; G.sph = nullptr; G.sph_k = INFINITY;
    const int nblocks = (int)((ncells + SCAN_BLOCK - 1) / SCAN_BLOCK);
    if (grow_buf((void **)&ctx->gb_count, &ctx->gb_cap[0], ncells * 4, "alloc grid count")) return 1;
    if (grow_buf((void **)&ctx->gb_raw_start, &ctx->gb_cap[1], (ncells + 1) * 4, "alloc grid start")) return 1;
    if (grow_buf((void **)&ctx->gb_cursor, &ctx->gb_cap[2], ncells * 4, "alloc grid cursor")) return 1;
    if (grow_buf((void **)&ctx->gb_bsums, &ctx->gb_cap[3], (size_t)(nblocks + 1) * 4, "alloc scan sums")) return 1;
    if (grow_buf((void **)&ctx->d_vlp_cell_start, &ctx->vlp_start_cap, (ncells + 1) * 4, "alloc VLP cell_start")) return 1;
    PT_CUDA(cudaMemsetAsync(ctx->gb_count, 0, ncells * 4, ctx->stream), "memset");
    PT_CUDA(cudaMemsetAsync(ctx->gb_cursor, 0, ncells * 4, ctx->stream), "memset");
    const int tb = 256, tg = (n + tb - 1) / tb;
    if (n > 0) {
        k_vlp_count<<<tg, tb, 0, ctx->stream>>>(ctx->d_vpls, n, G, ctx->gb_count);
        PT_CUDA(cudaGetLastError(), "VLP grid count");
    }
    if (exclusive_scan(ctx, ctx->gb_count, ncells, 0, ctx->gb_raw_start, ctx->gb_bsums)) return 1;
    if (exclusive_scan(ctx, ctx->gb_count, ncells, cap, ctx->d_vlp_cell_start, ctx->gb_bsums)) return 1;
    uint32_t raw_total = 0, cap_total = 0;
    PT_CUDA(cudaMemcpyAsync(&raw_total, ctx->gb_raw_start + ncells, 4, cudaMemcpyDeviceToHost, ctx->stream), "read total");
    PT_CUDA(cudaMemcpyAsync(&cap_total, ctx->d_vlp_cell_start + ncells, 4, cudaMemcpyDeviceToHost, ctx->stream), "read total");
    PT_CUDA(cudaStreamSynchronize(ctx->stream), "sync VLP grid totals");
    if (grow_buf((void **)&ctx->gb_raw_refs, &ctx->gb_cap[6], (size_t)(raw_total ? raw_total : 1) * 4, "alloc raw refs")) return 1;
    if (grow_buf((void **)&ctx->d_vlp_refs, &ctx->vlp_refs_cap, (size_t)(cap_total ? cap_total : 1) * 4, "alloc VLP refs")) return 1;
    if (n > 0) {
        k_vlp_fill<<<tg, tb, 0, ctx->stream>>>(ctx->d_vpls, n, G, ctx->gb_raw_start, ctx->gb_cursor, ctx->gb_raw_refs);
        PT_CUDA(cudaGetLastError(), "VLP grid fill");
    }
    k_vlp_emit<<<(unsigned)((ncells + 127) / 128), 128, 0, ctx->stream>>>(ncells, cap, ctx->gb_raw_start, ctx->gb_raw_refs,
                                                                       ctx->d_vlp_cell_start, ctx->d_vlp_refs);
    PT_CUDA(cudaGetLastError(), "VLP grid emit");
    ctx->vlp_grid_desc = *g;
    ctx->vlp_ncells = ncells;
    ctx->vlp_total_refs = cap_total;
    ctx->vlp_grid_set = true;
    return 0;
}

namespace pt {
// cells -> cells inside a one-cell border of sentinel words (GridDev::cells_pad); one thread per padded cell
__global__ void k_grid_pad(const uint2 *__restrict__ cells, int rx, int ry, int rz, uint2 *__restrict__ pad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t sx = rx + 2, sy = ry + 2, n = sx * sy * (size_t)(rz + 2);
    if (i >= n) return;
    const int x = (int)(i % sx) - 1, y = (int)((i / sx) % sy) - 1, z = (int)(i / (sx * sy)) - 1;
    const bool in = x >= 0 && x < rx && y >= 0 && y < ry && z >= 0 && z < rz;
    pad[i] = in ? cells[((size_t)z * ry + y) * rx + x] : make_uint2(0u, 0xFFFFFFFFu);
}
}  // namespace pt

int pt_grid_build_device(pt_ctx ctx, const pt_grid *g) {
    using namespace pt;
    const size_t ncells = (size_t)g->res[0] * g->res[1] * g->res[2];
    const uint32_t cap = g->max_refs_per_cell > 0 ? (uint32_t)g->max_refs_per_cell : 62u;
    const int ntri = ctx->ntri_total;
    GridDev G;
    for (int a = 0; a < 3; ++a) {
        G.bmin[a] = g->box_min[a]; G.bmax[a] = g->box_max[a]; G.cell[a] = g->cell_size[a]; G.res[a] = g->res[a];
    }
    G.cells = nullptr; G.cells_pad = nullptr; G.pad_sx = G.res[0] + 2; G.pad_sxy = G.pad_sx * (G.res[1] + 2); G.recs = nullptr; G.sph = nullptr; G.sph_k = INFINITY;
    const size_t npad = (size_t)(g->res[0] + 2) * (g->res[1] + 2) * (g->res[2] + 2);
    if (g->res[0] < 1 || g->res[1] < 1 || g->res[2] < 1 || npad >= (1ull << 31))
        return pt_fail(1, "pt_build_grid: grid resolution %d x %d x %d out of range", g->res[0], g->res[1], g->res[2]);

    // All buffers of the build live in the context and only ever grow: cudaMalloc / cudaFree of the ~250 MB a 1 M-
    // triangle grid needs cost 80 ms per build, 300x the 0.24 ms the nine kernels take.
    const int nblocks = (int)((ncells + SCAN_BLOCK - 1) / SCAN_BLOCK);
    auto grow = [&](void **ptr, size_t *capacity, size_t bytes, const char *what) -> int {
        if (*ptr && *capacity >= bytes) return 0;
        if (*ptr) cudaFree(*ptr);
        *ptr = nullptr;
        *capacity = 0;
        PT_CUDA(cudaMalloc(ptr, bytes), what);
        *capacity = bytes;
        return 0;
    };
    if (grow((void **)&ctx->gb_count, &ctx->gb_cap[0], ncells * 4, "alloc grid count")) return 1;
    if (grow((void **)&ctx->gb_raw_start, &ctx->gb_cap[1], (ncells + 1) * 4, "alloc grid start")) return 1;
    if (grow((void **)&ctx->gb_cursor, &ctx->gb_cap[2], ncells * 4, "alloc grid cursor")) return 1;
    if (grow((void **)&ctx->gb_bsums, &ctx->gb_cap[3], (size_t)(nblocks + 1) * 4, "alloc scan sums")) return 1;
    if (grow((void **)&ctx->d_cell_start, &ctx->gb_cap[4], (ncells + 1) * 4, "alloc cell_start")) return 1;
    if (grow((void **)&ctx->d_cells, &ctx->gb_cap[5], ncells * sizeof(uint2), "alloc cells")) return 1;
    if (grow((void **)&ctx->d_cells_pad, &ctx->gb_cap[11], npad * sizeof(uint2), "alloc padded cells")) return 1;
    uint32_t *d_count = ctx->gb_count, *d_raw_start = ctx->gb_raw_start, *d_cursor = ctx->gb_cursor, *d_bsums = ctx->gb_bsums;
    PT_CUDA(cudaMemsetAsync(d_count, 0, ncells * 4, ctx->stream), "memset");
    PT_CUDA(cudaMemsetAsync(d_cursor, 0, ncells * 4, ctx->stream), "memset");
    if (grow((void **)&ctx->gb_kmax, &ctx->gb_cap[9], 4, "alloc kmax")) return 1;
    PT_CUDA(cudaMemsetAsync(ctx->gb_kmax, 0, 4, ctx->stream), "memset");

    const int tb = 256, tg = (ntri + tb - 1) / tb;
    uint32_t raw_total = 0, cap_total = 0, kmax_bits = 0;
    if (ntri > 0) {
        k_grid_count<<<tg, tb, 0, ctx->stream>>>(ctx->d_tris_raw, ntri, G, d_count, ctx->gb_kmax);
        PT_CUDA(cudaGetLastError(), "grid count");
    }
    if (exclusive_scan(ctx, d_count, ncells, 0, d_raw_start, d_bsums)) return 1;
    if (exclusive_scan(ctx, d_count, ncells, cap, ctx->d_cell_start, d_bsums)) return 1;
    PT_CUDA(cudaMemcpyAsync(&raw_total, d_raw_start + ncells, 4, cudaMemcpyDeviceToHost, ctx->stream), "read total");
    PT_CUDA(cudaMemcpyAsync(&cap_total, ctx->d_cell_start + ncells, 4, cudaMemcpyDeviceToHost, ctx->stream), "read total");
    PT_CUDA(cudaMemcpyAsync(&kmax_bits, ctx->gb_kmax, 4, cudaMemcpyDeviceToHost, ctx->stream), "read kmax");
    PT_CUDA(cudaStreamSynchronize(ctx->stream), "sync grid totals");
    if (grow((void **)&ctx->gb_raw_refs, &ctx->gb_cap[6], (size_t)(raw_total ? raw_total : 1) * 4, "alloc raw refs")) return 1;
    if (grow((void **)&ctx->d_refs, &ctx->gb_cap[7], (size_t)(cap_total ? cap_total : 1) * 4, "alloc refs")) return 1;
    if (grow((void **)&ctx->d_recs, &ctx->gb_cap[8], (size_t)(cap_total ? cap_total : 1) * 3 * sizeof(float4), "alloc records")) return 1;
    if (grow((void **)&ctx->d_sph, &ctx->gb_cap[10], (size_t)(cap_total ? cap_total : 1) * sizeof(float4), "alloc record spheres")) return 1;
    uint32_t *d_raw_refs = ctx->gb_raw_refs;
    if (ntri > 0) {
        k_grid_fill<<<tg, tb, 0, ctx->stream>>>(ctx->d_tris_raw, ntri, G, d_raw_start, d_cursor, d_raw_refs);
        PT_CUDA(cudaGetLastError(), "grid fill");
    }
    k_grid_emit<<<(unsigned)((ncells + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_tris_raw, ncells, cap, d_raw_start, d_raw_refs,
                                                                        ctx->d_cell_start, ctx->d_refs, ctx->d_recs, ctx->d_sph,
                                                                        ctx->d_cells);
    PT_CUDA(cudaGetLastError(), "grid emit");
    k_grid_pad<<<(unsigned)((npad + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_cells, g->res[0], g->res[1], g->res[2], ctx->d_cells_pad);
    PT_CUDA(cudaGetLastError(), "grid pad");
    // no synchronisation here: the render that follows is ordered behind the build on the same stream

    G.cells = ctx->d_cells;
    G.cells_pad = ctx->d_cells_pad;
    G.recs = ctx->d_recs;
    G.sph = ctx->d_sph;
    {
        float kmax;
        memcpy(&kmax, &kmax_bits, 4);
        const double k = 2e-4 * (double)kmax + 1e-6;          // same margin law as the brute-force mesh cull (pt_set_scene)
        G.sph_k = (kmax_bits < 0x7f800000u && k < 1e30) ? (float)k : INFINITY;
    }
    ctx->grid = G;
    ctx->grid_desc = *g;
    ctx->ncells = ncells;
    ctx->total_refs = cap_total;
    ctx->grid_set = true;
    return 0;
}
