// pt_metropolis.cuh — kernels lightTracer (seed paths) and MetropolisLightTracer of CLSuperMetropolisPathTracer(_vlpgrid)/
// metropolispathtracer.ocl, in FIX mode.
//
// As written those kernels have no defined behaviour: VerifyIntersection hands TraceRay an UNINITIALISED `float t` as its
// running hit bound (:225-236, vlpgrid :239-242) and the host passes the VPL buffer where lightTracer expects the seed-path
// buffer (CLSuperMetropolisPathTracer.c:439 / vlpgrid :579).  FIX mode resolves exactly these two points — t = 1e9, the
// default distance every other caller of TraceRay uses; seed paths in their own buffer — and reproduces everything else as
// written, including that every helper takes the RNG state BY VALUE (:146,159,172,184,241): a work-item's "random"
// directions repeat, GetRandomPath shoots its four segments along one direction, every mutation round draws the same numbers.
// The checker is the reference itself compiled with that one-line patch (oracle/Makefile, libref_vlpgrid_fix.so).
// One thread per seed path (512 work-items per light by default): nothing here is performance-critical.
#pragma once
#include "pt_bidir.cuh"

namespace pt {

struct MPath { V3 v[4]; uint32_t length; };

template <bool FMA>
struct Metro {
    typedef Ar<FMA> A;
    const LaunchArgs &P;
    const SceneBlock *S;
    Counters cnt;
    PT_DEV Metro(const LaunchArgs &p) : P(p), S(p.gscene) { cnt = {0, 0, 0, 0, 0, 0}; }

    // TraceRay (lmem family: running bound t) -> material, t, normal
    PT_DEV int trace(V3 o, V3 d, float &t, V3 &n) {
        const int hit = trace_ray<FMA, true, false>(P.ap, S, P.grid, o, d, t, cnt);
        if (hit == HIT_NONE) return 0;
        n = hit_normal<FMA, false>(P.ap, S, P.grid, hit, o, d, t);
        return hit_material(hit);
    }
    // :146-156 — rng by value
    PT_DEV V3 random_direction(Rng rng) {
        float r0 = 0.0f, r1 = 0.0f, sum = 2.0f;
        while (sum >= 1.0f) {
            rng_next_pm1(rng, r0, r1);
            sum = A::madd(r1, r1, A::mul(r0, r0));
        }
        const float sq = A::sqrt(A::sub(1.0f, sum));
        return mk3(A::mul(A::mul(2.0f, r0), sq), A::mul(A::mul(2.0f, r1), sq), A::sub(1.0f, A::mul(2.0f, sum)));
    }
    // :158-170
    PT_DEV bool add_vertex(V3 origin, V3 &vertex, MPath &path, Rng rng) {
        const V3 d = random_direction(rng);
        V3 n = mk3(0.f, 0.f, 0.f);
        float t = 1e9f;
        if (trace(origin, d, t, n)) {
            vertex = A::vmadd(d, t, origin);
            path.length += 1;
            return true;
        }
        return false;
    }
    // :172-182
    PT_DEV MPath random_path(V3 origin, Rng rng) {
        MPath p;
        p.length = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) p.v[i] = mk3(0.f, 0.f, 0.f);
        V3 cur = origin;
        if (!add_vertex(cur, p.v[0], p, rng)) return p;
        cur = p.v[0];
        if (!add_vertex(cur, p.v[1], p, rng)) return p;
        cur = p.v[1];
        if (!add_vertex(cur, p.v[2], p, rng)) return p;
        cur = p.v[2];
        add_vertex(cur, p.v[3], p, rng);
        return p;
    }
    static PT_DEV float perturb1(float vertex, float r, float dx) {
        if (r < 0.5f) return vertex < 1.0f ? A::add(vertex, dx) : A::sub(A::add(vertex, dx), 1.0f);
        return vertex < 0.0f ? A::add(A::sub(vertex, dx), 1.0f) : A::sub(vertex, dx);
    }
    // :184-221 — two RNG pairs (x, y from the first, z from the second), rng by value; every operation rounded on its own
    PT_DEV V3 perturbation(V3 vertex, Rng rng) {
        float a0, a1, b0, b1;
        rng_next(rng, a0, a1);
        rng_next(rng, b0, b1);
        const float s1 = 1.0f / 512.0f, s2 = 1.0f / 16.0f;
        const float q = A::div(s1, s2), tail = A::div(s1, A::add(q, 1.0f));
        const float dx = A::sub(A::div(s1, A::add(q, fabsf(A::sub(A::mul(2.0f, a0), 1.0f)))), tail);
        const float dy = A::sub(A::div(s1, A::add(q, fabsf(A::sub(A::mul(2.0f, a1), 1.0f)))), tail);
        const float dz = A::sub(A::div(s1, A::add(q, fabsf(A::sub(A::mul(2.0f, b0), 1.0f)))), tail);
        return mk3(perturb1(vertex.x, a0, dx), perturb1(vertex.y, a1, dy), perturb1(vertex.z, b0, dz));
    }
    // :223-236 with the FIX: t = 1e9
    PT_DEV bool verify(V3 origin, V3 dest) {
        float t = 1e9f;
        V3 n = mk3(0.f, 0.f, 0.f);
        const V3 d = A::normalize(A::vsub(dest, origin));
        if (!trace(origin, d, t, n)) return false;
        const V3 X = A::vmadd(d, t, origin);
        return dest.x == X.x && dest.y == X.y && dest.z == X.z;
    }
    // :238-294
    PT_DEV void mutate(MPath &seed, V3 origin, Rng rng) {
        if (seed.length == 0) {
            seed = random_path(origin, rng);
            if (seed.length == 0) return;
        }
        float y0, y1;
        rng_next(rng, y0, y1);
        const float prob = A::div(1.0f, A::add(__uint2float_rn(seed.length), 0.2f));
        if (prob < y0) return;
        MPath tmp;
        tmp.length = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) tmp.v[i] = mk3(0.f, 0.f, 0.f);
        V3 cur = origin;
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) {
            if (i >= seed.length) break;
            tmp.v[i] = perturbation(seed.v[i], rng);
            if (verify(cur, tmp.v[i])) { tmp.length++; cur = tmp.v[i]; }
            else break;
        }
        if (tmp.length == seed.length) seed = tmp;
        if (seed.length == 1) {
            if (y1 > 0.3f) { if (!add_vertex(seed.v[0], seed.v[1], seed, rng)) return; }
            if (y1 > 0.7f) { if (!add_vertex(seed.v[1], seed.v[2], seed, rng)) return; }
            if (y1 > 0.9f) add_vertex(seed.v[2], seed.v[3], seed, rng);
        } else if (seed.length == 2) {
            if (y1 < 0.3f) { if (!add_vertex(seed.v[1], seed.v[2], seed, rng)) return; }
            if (y1 < 0.2f) add_vertex(seed.v[2], seed.v[3], seed, rng);
        } else if (seed.length == 3) {
            if (y1 < 0.2f) add_vertex(seed.v[2], seed.v[3], seed, rng);
        }
    }
    // :380-428 SampleFromLightSource of the Metropolis programs (constants 400 / 10 / 40, total_paths / 256)
    PT_DEV float4 sample_from_light(V3 o, V3 d, float I, float denom) {
        float t = 1e9f;
        V3 n = mk3(0.f, 0.f, 0.f);
        const int m = trace(o, d, t, n);
        if (!m) return make_float4(0.f, 0.f, 0.f, 0.f);
        const V3 X = A::vmadd(d, t, o);
        float lam = A::dot(d, n);
        if (lam < 0.0f) lam = 0.0f;
        else {
            const V3 dv = A::vsub(o, X);
            const float dist = A::sqrt(A::dot(dv, dv));
            float f = A::div(I, A::mul(dist, dist));
            f = 1.0f < f ? 1.0f : f;
            lam = A::mul(lam, f);
        }
        if (lam > 1.0f) lam = 1.0f;
        const float k = m == 1 ? 400.0f : (m == 3 ? 40.0f : 0.0f);      // material 2 is never produced by TraceRay
        if (k == 0.0f) return make_float4(0.f, 0.f, 0.f, 0.f);
        return make_float4(X.x, X.y, X.z, A::div(A::mul(k, lam), denom));
    }
};

PT_DEV void store_path(uint32_t *q, const MPath &p) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        q[4 * i] = __float_as_uint(p.v[i].x); q[4 * i + 1] = __float_as_uint(p.v[i].y); q[4 * i + 2] = __float_as_uint(p.v[i].z); q[4 * i + 3] = 0u;
    }
    q[16] = p.length; q[17] = q[18] = q[19] = 0u;
}

// Both kernels of the reference in one launch (same work-items, same seeds: CLSuperMetropolisPathTracer.c:436-441).
// seed_paths / mutated_paths: n*nlights x 20 words in the reference's Path layout {float4 v[4]; uint length; pad[3]};
// vpl_out: 4*n*nlights float4, entry 4*(gi + l*n) + i.
template <bool FMA>
__global__ void __launch_bounds__(64) k_metropolis(const __grid_constant__ LaunchArgs P, int n, int rounds, float4 *__restrict__ vpl_out,
                                                   uint32_t *__restrict__ seed_paths, uint32_t *__restrict__ mutated_paths) {
    typedef Ar<FMA> A;
    const int gi = blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= n) return;
    Metro<FMA> M(P);
    const int nl = P.ap.nlights;
    const int total_paths = n * nl;
    const float denom = __int2float_rn(total_paths / 256);             // integer division first
    const Rng rng = rng_seed(P.seeds, (uint32_t)gi);
    for (int l = 0; l < nl; ++l) {
        const float4 L = P.ap.lights[l];
        V3 origin = mk3(L.x, L.y, L.z);
        MPath seed = M.random_path(origin, rng);                       // kernel lightTracer
        const size_t slot = (size_t)gi + (size_t)l * n;
        store_path(seed_paths + 20 * slot, seed);
        for (int m = 0; m < rounds; ++m) M.mutate(seed, origin, rng);   // kernel MetropolisLightTracer
        store_path(mutated_paths + 20 * slot, seed);
        float4 out[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if ((uint32_t)i >= seed.length) break;
            const V3 d = A::normalize(A::vsub(seed.v[i], origin));
            out[i] = M.sample_from_light(origin, d, A::div(L.w, __int2float_rn(1 << i)), denom);
            if (out[i].w == 0.0f) break;
            origin = seed.v[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) vpl_out[4 * slot + i] = out[i];
    }
}

}  // namespace pt

// seed paths + Metropolis pass + compaction of the resulting VPL buffer (vpl: 4*n*nlights entries)
int pt_launch_metropolis_kernels(pt_ctx ctx, int arith, const pt::LaunchArgs &args, int n, int rounds, float4 *vpl, uint32_t *seed_paths,
                                 uint32_t *mutated_paths, float4 *active, int *count) {
    using namespace pt;
    const int nl = args.ap.nlights;
    if (n > 0 && nl > 0) {
        if (arith != PT_ARITH_SEPARATE) k_metropolis<true><<<(n + 63) / 64, 64, 0, ctx->stream>>>(args, n, rounds, vpl, seed_paths, mutated_paths);
        else k_metropolis<false><<<(n + 63) / 64, 64, 0, ctx->stream>>>(args, n, rounds, vpl, seed_paths, mutated_paths);
        PT_CUDA(cudaGetLastError(), "launch k_metropolis");
    }
    k_compact_vpls<<<1, 256, 0, ctx->stream>>>(vpl, 4 * n * nl, active, count);
    PT_CUDA(cudaGetLastError(), "launch k_compact_vpls");
    return 0;
}
