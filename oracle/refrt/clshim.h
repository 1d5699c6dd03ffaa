/*
 * clshim.h — just enough of the OpenCL C 1.2 *device* language, expressed in
 * C++17, to compile the reference's own pathtracer.ocl files with g++.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/_ref): the product path never includes it.
 *
 * The reference kernels are compiled from where they lie under
 * /root/reference (never copied into this repo): oracle/refrt/ocl2cpp.py
 * rewrites the one construct C++ cannot parse — OpenCL vector literals
 * `(float4)(a, b, c, d)` — into brace-initialisation `float4{a, b, c, d}`
 * (which, unlike a function call, also pins left-to-right evaluation of the two
 * MWC64XVEC2() calls in `pathtracer.ocl:233`), and the result is #included
 * inside `namespace ocl { ... }` after this header.
 *
 * Semantics follow the OpenCL 1.2 specification:
 *   - vector relational operators yield -1 (all bits set) per true lane
 *     (spec 6.3.d) — this is what makes `c = hi + convert_uint2(x < c)`
 *     (pathtracer.ocl:19) add 0xFFFFFFFF on carry;
 *   - scalar operands of vector operators are converted to the element type;
 *   - convert_T() without _sat/_rt* truncates toward zero; out-of-range input
 *     is implementation-defined, here it saturates (NaN -> 0) like NVIDIA GPUs;
 *   - dot() is summed left to right over x,y,z,w; every operation is a separate
 *     IEEE-754 binary32 operation (no contraction: build without -mfma).
 */
#ifndef REFRT_CLSHIM_H
#define REFRT_CLSHIM_H

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <type_traits>

namespace ocl {

typedef unsigned int   uint;
typedef unsigned short ushort;
typedef unsigned char  uchar;

template <class S>
struct is_scalar : std::integral_constant<bool, std::is_arithmetic<S>::value || std::is_enum<S>::value> {};

/* ------------------------------------------------------------------ vectors */
template <class T>
struct vec2 {
    union {
        struct { T x, y; };
        struct { T s0, s1; };
        T s[2];
    };
    vec2() = default;
    template <class A, class B, class = typename std::enable_if<is_scalar<A>::value && is_scalar<B>::value>::type>
    vec2(A a, B b) { x = (T)a; y = (T)b; }
    template <class A, class = typename std::enable_if<is_scalar<A>::value>::type>
    explicit vec2(A a) { x = (T)a; y = (T)a; }
};

/* 3-component swizzle result (`v.s012`, rewritten to `v.s012()` by ocl2cpp.py) */
template <class T>
struct vec3 { T x, y, z; };

template <class T>
struct alignas(sizeof(T) * 4) vec4 {
    union {
        struct { T x, y, z, w; };
        struct { T s0, s1, s2, s3; };
        T s[4];
    };
    vec4() = default;
    template <class A, class B, class C, class D,
              class = typename std::enable_if<is_scalar<A>::value && is_scalar<B>::value &&
                                              is_scalar<C>::value && is_scalar<D>::value>::type>
    vec4(A a, B b, C c, D d) { x = (T)a; y = (T)b; z = (T)c; w = (T)d; }
    template <class C, class D, class = typename std::enable_if<is_scalar<C>::value && is_scalar<D>::value>::type>
    vec4(vec2<T> ab, C c, D d) { x = ab.x; y = ab.y; z = (T)c; w = (T)d; }
    vec4(vec2<T> ab, vec2<T> cd) { x = ab.x; y = ab.y; z = cd.x; w = cd.y; }
    template <class D, class = typename std::enable_if<is_scalar<D>::value>::type>
    vec4(vec3<T> abc, D d) { x = abc.x; y = abc.y; z = abc.z; w = (T)d; }
    vec3<T> s012() const { vec3<T> r; r.x = x; r.y = y; r.z = z; return r; }
    template <class A, class = typename std::enable_if<is_scalar<A>::value>::type>
    explicit vec4(A a) { x = (T)a; y = (T)a; z = (T)a; w = (T)a; }
};

/* 8- and 16-component float vectors: only what the VLP-grid kernels of CLSuperMetropolisPathTracer_vlpgrid use —
 * (float8)(lo, hi), .lo / .hi, whole-vector loads and stores */
struct alignas(64) float16 { float v[16]; };

typedef vec2<float> float2;
typedef vec2<uint>  uint2;
typedef vec2<int>   int2;
typedef vec4<float> float4;
typedef vec4<int>   int4;
typedef vec4<uint>  uint4;
typedef vec4<uchar> uchar4;

/* float8 as a pair of float4 (defined after float4 is complete) */
struct alignas(32) float8 {
    float4 lo, hi;
    float8() = default;
    float8(float4 a, float4 b) { lo = a; hi = b; }
};

#define REFRT_BINOP(OP)                                                                            \
    template <class T> inline vec4<T> operator OP(vec4<T> a, vec4<T> b) {                          \
        vec4<T> r; r.x = a.x OP b.x; r.y = a.y OP b.y; r.z = a.z OP b.z; r.w = a.w OP b.w; return r; } \
    template <class T, class S, class = typename std::enable_if<is_scalar<S>::value>::type>        \
    inline vec4<T> operator OP(vec4<T> a, S s) { return a OP vec4<T>((T)s); }                      \
    template <class T, class S, class = typename std::enable_if<is_scalar<S>::value>::type>        \
    inline vec4<T> operator OP(S s, vec4<T> b) { return vec4<T>((T)s) OP b; }                      \
    template <class T> inline vec2<T> operator OP(vec2<T> a, vec2<T> b) {                          \
        vec2<T> r; r.x = a.x OP b.x; r.y = a.y OP b.y; return r; }                                 \
    template <class T, class S, class = typename std::enable_if<is_scalar<S>::value>::type>        \
    inline vec2<T> operator OP(vec2<T> a, S s) { return a OP vec2<T>((T)s); }                      \
    template <class T, class S, class = typename std::enable_if<is_scalar<S>::value>::type>        \
    inline vec2<T> operator OP(S s, vec2<T> b) { return vec2<T>((T)s) OP b; }
REFRT_BINOP(+)
REFRT_BINOP(-)
REFRT_BINOP(*)
REFRT_BINOP(/)
REFRT_BINOP(^)
REFRT_BINOP(&)
#undef REFRT_BINOP

template <class T> inline vec4<T> operator-(vec4<T> a) { vec4<T> r; r.x = -a.x; r.y = -a.y; r.z = -a.z; r.w = -a.w; return r; }
template <class T, class U> inline vec4<T> &operator+=(vec4<T> &a, U b) { a = a + b; return a; }
template <class T, class U> inline vec4<T> &operator-=(vec4<T> &a, U b) { a = a - b; return a; }
template <class T, class U> inline vec4<T> &operator*=(vec4<T> &a, U b) { a = a * b; return a; }

/* shifts: vector << scalar int */
inline int4 operator<<(int4 a, int n) {
    int4 r;
    r.x = (int)((uint)a.x << n); r.y = (int)((uint)a.y << n);
    r.z = (int)((uint)a.z << n); r.w = (int)((uint)a.w << n);
    return r;
}

/* vector relational operators: -1 per true lane (OpenCL 1.2, 6.3.d) */
template <class T> inline int2 operator<(vec2<T> a, vec2<T> b) { return int2(a.x < b.x ? -1 : 0, a.y < b.y ? -1 : 0); }

/* -------------------------------------------------------------- conversions */
inline uint2  convert_uint2(int2 a)   { return uint2((uint)a.x, (uint)a.y); }
inline float2 convert_float2(uint2 a) { return float2((float)a.x, (float)a.y); }  /* round-to-nearest-even */
inline float4 convert_float4(int4 a)  { return float4((float)a.x, (float)a.y, (float)a.z, (float)a.w); }
inline int refrt_f2i_rz_sat(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (-2147483647 - 1);
    return (int)f;
}
inline uchar refrt_f2u8_rz_sat(float f) {
    if (f != f) return 0;
    if (f >= 255.0f) return 255;
    if (f <= 0.0f) return 0;
    return (uchar)(int)f;
}
inline int4 convert_int4(float4 a) {
    return int4(refrt_f2i_rz_sat(a.x), refrt_f2i_rz_sat(a.y), refrt_f2i_rz_sat(a.z), refrt_f2i_rz_sat(a.w));
}
inline uchar4 convert_uchar4(float4 a) {
    return uchar4(refrt_f2u8_rz_sat(a.x), refrt_f2u8_rz_sat(a.y), refrt_f2u8_rz_sat(a.z), refrt_f2u8_rz_sat(a.w));
}

/* --------------------------------------------------------------------- math */
inline float sqrt(float x) { return ::sqrtf(x); }
inline float fabs(float x) { return ::fabsf(x); }
inline float4 fabs(float4 a) { return float4(::fabsf(a.x), ::fabsf(a.y), ::fabsf(a.z), ::fabsf(a.w)); }
inline float floor(float x) { return ::floorf(x); }
inline float ceil(float x) { return ::ceilf(x); }
template <class E, class = typename std::enable_if<is_scalar<E>::value>::type>
inline float pow(float x, E e) { return ::powf(x, (float)e); }

/* OpenCL min/max(x,y): "y < x ? y : x" / "x < y ? y : x" (spec 6.12.4) */
inline float min(float x, float y) { return y < x ? y : x; }
inline float max(float x, float y) { return x < y ? y : x; }
inline int   min(int x, int y) { return y < x ? y : x; }
inline int   max(int x, int y) { return x < y ? y : x; }
/* fmin/fmax: NaN-ignoring */
inline float fmin(float x, float y) { if (x != x) return y; if (y != y) return x; return y < x ? y : x; }
inline float fmax(float x, float y) { if (x != x) return y; if (y != y) return x; return x < y ? y : x; }
inline float4 fmin(float4 a, float4 b) { return float4(fmin(a.x, b.x), fmin(a.y, b.y), fmin(a.z, b.z), fmin(a.w, b.w)); }
inline float4 fmax(float4 a, float4 b) { return float4(fmax(a.x, b.x), fmax(a.y, b.y), fmax(a.z, b.z), fmax(a.w, b.w)); }
inline int clamp(int v, int lo, int hi) { return min(max(v, lo), hi); }
inline int4 clamp(int4 v, int4 lo, int4 hi) {
    return int4(clamp(v.x, lo.x, hi.x), clamp(v.y, lo.y, hi.y), clamp(v.z, lo.z, hi.z), clamp(v.w, lo.w, hi.w));
}

inline float dot(float4 a, float4 b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
inline float4 cross(float4 a, float4 b) {
    return float4(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x, 0.0f);
}
inline float length(float4 a) { return sqrt(dot(a, a)); }
inline float distance(float4 a, float4 b) { return length(a - b); }

template <class S, class = typename std::enable_if<is_scalar<S>::value>::type>
inline int4 isgreater(float4 a, S b) {
    float f = (float)b;
    return int4(a.x > f ? -1 : 0, a.y > f ? -1 : 0, a.z > f ? -1 : 0, a.w > f ? -1 : 0);
}
template <class S, class = typename std::enable_if<is_scalar<S>::value>::type>
inline int4 isgreaterequal(float4 a, S b) {
    float f = (float)b;
    return int4(a.x >= f ? -1 : 0, a.y >= f ? -1 : 0, a.z >= f ? -1 : 0, a.w >= f ? -1 : 0);
}
inline int4 isless(float4 a, float4 b) { return int4(a.x < b.x ? -1 : 0, a.y < b.y ? -1 : 0, a.z < b.z ? -1 : 0, a.w < b.w ? -1 : 0); }
inline int4 isgreater(float4 a, float4 b) { return int4(a.x > b.x ? -1 : 0, a.y > b.y ? -1 : 0, a.z > b.z ? -1 : 0, a.w > b.w ? -1 : 0); }
/* select(a, b, c): per lane, MSB(c) ? b : a */
template <class T> inline vec4<T> select(vec4<T> a, vec4<T> b, int4 c) {
    return vec4<T>(c.x < 0 ? b.x : a.x, c.y < 0 ? b.y : a.y, c.z < 0 ? b.z : a.z, c.w < 0 ? b.w : a.w);
}

inline uint mul_hi(uint a, uint b) { return (uint)(((uint64_t)a * (uint64_t)b) >> 32); }
template <class S> inline uint2 mul_hi(uint2 a, S b) { return uint2(mul_hi(a.x, (uint)b), mul_hi(a.y, (uint)b)); }

inline int atomic_inc(volatile uint *p) { return (int)__atomic_fetch_add(p, 1u, __ATOMIC_RELAXED); }
inline int atomic_inc(volatile int *p)  { return __atomic_fetch_add(p, 1, __ATOMIC_RELAXED); }

/* ------------------------------------------------------- work-item functions */
struct WorkItem {
    size_t gid[3], lid[3], grp[3];
    const size_t *gsz, *lsz, *ngrp;
};
extern thread_local WorkItem *refrt_wi;
void refrt_barrier();

inline size_t get_global_id(uint d)   { return refrt_wi->gid[d]; }
inline size_t get_local_id(uint d)    { return refrt_wi->lid[d]; }
inline size_t get_group_id(uint d)    { return refrt_wi->grp[d]; }
inline size_t get_global_size(uint d) { return refrt_wi->gsz[d]; }
inline size_t get_local_size(uint d)  { return refrt_wi->lsz[d]; }
inline size_t get_num_groups(uint d)  { return refrt_wi->ngrp[d]; }
enum { CLK_LOCAL_MEM_FENCE = 1, CLK_GLOBAL_MEM_FENCE = 2 };
inline void barrier(int) { refrt_barrier(); }

}  // namespace ocl

/* address-space / function qualifiers: meaningless on a CPU.  Defined LAST so
 * they cannot disturb the standard headers above. */
#define kernel
#define global
#define local
#define constant
#define restrict __restrict__

#endif /* REFRT_CLSHIM_H */
