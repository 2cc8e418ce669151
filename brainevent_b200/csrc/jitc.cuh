// jitc.cuh -- light-RNG device functions and the JITC parameter block, shared by jitc.cu (products)
// and jitc_csr.cu (materialisation).  RNG spec: ref _numba_random.py:385-502.
#pragma once
#include "common.cuh"

namespace be {

// ------------------------------------------------------------------------------------------
// light-RNG (all uint32 wrap-around arithmetic)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// xorshift32 (13,17,5) is a bijection on non-zero states, so the reference's `x == 0` guard can
// only fire for a zero input, which init() already excludes; it is kept for exactness of intent.
__device__ __forceinline__ uint32_t rng_next(uint32_t x) {
    x ^= x << 13; x ^= x >> 17; x ^= x << 5;
    return x;
}
__device__ __forceinline__ uint32_t rng_init(uint32_t seed, uint32_t row, uint32_t chunk, uint32_t lane) {
    uint32_t x = seed ^ 0xd1b54a35U;
    x ^= row * 0x85ebca6bU; x ^= chunk * 0xc2b2ae35U; x ^= lane * 0x27d4eb2dU;
    x = mix32(x);
    return x == 0U ? 0x6d2b79f5U : x;
}
__device__ __forceinline__ uint32_t rng_initial_q(uint32_t& s, uint32_t cl) {
    const uint32_t n = cl - 1U;
    while (true) {
        s = rng_next(s); const uint32_t q = __umulhi(s, n);
        s = rng_next(s); const uint32_t g = __umulhi(s, n);
        if (g < n - q) return q;
    }
}
__device__ __forceinline__ float hash_uniform01(uint32_t seed, uint32_t row, uint32_t col) {
    uint32_t h = seed ^ 0xa0761d65U;
    h ^= row * 0xe7037ed1U; h ^= col * 0x8ebc6af1U;
    h = mix32(h);
    return __fmul_rn((float)(h & 0x00ffffffU), 1.0f / 16777216.0f);
}
// Horner steps with explicit rounding (no FMA contraction): bit-identical to strict IEEE f32.
__device__ __forceinline__ float hs(float acc, float x, float c) { return __fadd_rn(__fmul_rn(acc, x), c); }
// Acklam's inverse normal CDF in three pieces so that callers can evaluate the common central branch
// inline and hand the rare tails (4.85 % of the draws; logf + sqrt + two more Horner chains) to
// converged lanes.  hash_normal01 composes them exactly as before: same operations, same roundings.
__device__ __forceinline__ float normal01_clamp_u(float u) { return fmaxf(fminf(u, 1.0f - 1e-10f), 1e-10f); }
__device__ __forceinline__ bool normal01_is_tail(float u) { return u < 0.02425f || u > 0.97575f; }
__device__ __forceinline__ float normal01_tail(float u) {
    const float c1 = -0.007784894f, c2 = -0.32239646f, c3 = -2.4007583f, c4 = -2.5497325f,
                c5 = 4.3746641f, c6 = 2.9381640f;
    const float d1 = 0.007784696f, d2 = 0.32246713f, d3 = 2.4451342f, d4 = 3.7544087f;
    const bool lower = u < 0.02425f;
    const float t = lower ? u : __fsub_rn(1.0f, u);
    const float v = __fsqrt_rn(__fmul_rn(-2.0f, logf(t)));
    const float num = hs(hs(hs(hs(hs(c1, v, c2), v, c3), v, c4), v, c5), v, c6);
    const float den = hs(hs(hs(hs(d1, v, d2), v, d3), v, d4), v, 1.0f);
    const float z = __fdiv_rn(num, den);
    return lower ? -z : z;
}
__device__ __forceinline__ float normal01_central(float u) {
    const float a1 = -39.696830f, a2 = 220.94609f, a3 = -275.92851f, a4 = 138.35775f,
                a5 = -30.664799f, a6 = 2.5066283f;
    const float b1 = -54.476099f, b2 = 161.58584f, b3 = -155.69898f, b4 = 66.801312f,
                b5 = -13.280681f;
    const float v = __fsub_rn(u, 0.5f);
    const float r = __fmul_rn(v, v);
    const float num = __fmul_rn(hs(hs(hs(hs(hs(a1, r, a2), r, a3), r, a4), r, a5), r, a6), v);
    const float den = hs(hs(hs(hs(hs(b1, r, b2), r, b3), r, b4), r, b5), r, 1.0f);
    return __fdiv_rn(num, den);
}
__device__ __forceinline__ float hash_normal01(uint32_t seed, uint32_t row, uint32_t col) {
    const float u = normal01_clamp_u(hash_uniform01(seed, row, col));
    return normal01_is_tail(u) ? normal01_tail(u) : normal01_central(u);
}

template <int KIND, typename A>
__device__ __forceinline__ A jit_weight(A p0, A p1, uint32_t seed, uint32_t row, uint32_t col) {
    if (KIND == BE_B200_JIT_SCALAR) return p0;
    if (KIND == BE_B200_JIT_NORMAL) return p0 + (A)hash_normal01(seed, row, col) * p1;
    return p0 + (A)hash_uniform01(seed, row, col) * p1;  // p1 holds (w_high - w_low)
}
// rounding note: the oracle computes w_loc + n01*w_scale in the weight dtype; for f32 the
// compiler may fuse the multiply-add here, which is within 1 ulp of the unfused value.

struct JitParams {
    const void* p0;
    const void* p1;
    const int32_t* clen;
    const int32_t* seed;
};
template <typename T, int KIND>
__device__ __forceinline__ void load_params(const JitParams& jp, typename Acc<T>::type& p0,
                                            typename Acc<T>::type& p1, uint32_t& cl, uint32_t& seed) {
    using A = typename Acc<T>::type;
    p0 = to_acc<T>(__ldg((const T*)jp.p0));
    p1 = A(0);
    if (KIND == BE_B200_JIT_NORMAL) p1 = to_acc<T>(__ldg((const T*)jp.p1));
    if (KIND == BE_B200_JIT_UNIFORM) p1 = to_acc<T>(__ldg((const T*)jp.p1)) - p0;
    cl = (uint32_t)__ldg(jp.clen);
    if (cl < 2U) cl = 2U;
    seed = (uint32_t)__ldg(jp.seed);
}

// number of walk positions of `lane` inside a chunk of `width` columns with residue stride S
template <int S>
__device__ __forceinline__ uint32_t lane_qmax(uint32_t width, uint32_t lane) {
    return width > lane ? (width - lane + (S - 1)) / S : 0U;
}

}  // namespace be

#define BE_DISPATCH_KIND(kind, K, ...)                                             \
    switch (kind) {                                                                \
        case BE_B200_JIT_SCALAR: { constexpr int K = BE_B200_JIT_SCALAR; __VA_ARGS__; break; }   \
        case BE_B200_JIT_NORMAL: { constexpr int K = BE_B200_JIT_NORMAL; __VA_ARGS__; break; }   \
        case BE_B200_JIT_UNIFORM: { constexpr int K = BE_B200_JIT_UNIFORM; __VA_ARGS__; break; } \
        default: ::be::set_error("unknown JITC kind %d", (int)(kind)); return BE_B200_EINVAL;    \
    }
