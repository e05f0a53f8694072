// Device restatement of rten-vecmath's scalar recipes.  Same operation
// order, fused multiply-adds and exact IEEE division as the reference's AVX-512 path, so the
// elementwise results are bit-identical to it:
//   ReducedRangeExp / Exp : rten-vecmath/src/exp.rs:61-191
//   Erf / Gelu / ApproxGelu: rten-vecmath/src/erf.rs:23-100
//   Tanh                  : rten-vecmath/src/tanh.rs:12-66
// Compile WITHOUT --use_fast_math (division and fmaf must stay IEEE).
#pragma once
#include <cstdint>

namespace rtb {

__device__ __forceinline__ float exp_poly(float x, float& j) {
    const float inv_log2 = 1.44269504088896340736f;
    const float magic = 12582912.0f;
    j = __fmaf_rn(x, inv_log2, magic);
    j = __fsub_rn(j, magic);
    float r = __fmaf_rn(j, -6.93145752e-1f, x);
    r = __fmaf_rn(j, -1.42860677e-6f, r);
    float t = 1.37805939e-3f;
    t = __fmaf_rn(t, r, 8.37312452e-3f);
    t = __fmaf_rn(t, r, 4.16695364e-2f);
    t = __fmaf_rn(t, r, 1.66664720e-1f);
    t = __fmaf_rn(t, r, 4.99999851e-1f);
    t = __fmaf_rn(t, r, 1.0f);
    return __fmaf_rn(t, r, 1.0f);
}

// x86 cvttps2dq semantics: NaN / out of range -> INT_MIN
__device__ __forceinline__ int trunc_i32_x86(float x) {
    if (!(x > -2147483904.0f && x < 2147483648.0f)) return (int)0x80000000;
    return __float2int_rz(x);
}

__device__ __forceinline__ float reduced_range_exp(float x) {
    const float cutoff = -126.5f * 0.693147180559945309417f + 0.01f;
    float j;
    float r = exp_poly(x, j);
    int k = trunc_i32_x86(j);
    float p2 = __int_as_float((int)((unsigned)(k + 127) << 23));
    r = __fmul_rn(r, p2);
    return (x < cutoff) ? 0.0f : r;
}

__device__ __forceinline__ float exp_ref(float x) {
    float j;
    float r = exp_poly(x, j);
    int k = trunc_i32_x86(j);
    unsigned ia = (k > 0) ? 0u : 0x83000000u;
    unsigned is = ia + 0x7f000000u;
    unsigned it = ((unsigned)k << 23) - ia;
    r = __fmul_rn(r, __uint_as_float(is));
    r = __fmul_rn(r, __uint_as_float(it));
    if (x >= 104.0f) r = __int_as_float(0x7f800000);
    if (x <= -104.0f) r = 0.0f;
    return r;
}

__device__ __forceinline__ float erf_ref(float x0) {
    bool neg = x0 < 0.0f;
    float x = fabsf(x0);
    float t = __fdiv_rn(1.0f, __fmaf_rn(x, 0.3275911f, 1.0f));
    float y = 1.061405429f;
    y = __fmaf_rn(y, t, -1.453152027f);
    y = __fmaf_rn(y, t, 1.421413741f);
    y = __fmaf_rn(y, t, -0.284496736f);
    y = __fmaf_rn(y, t, 0.254829592f);
    float at = __fmul_rn(y, t);
    float xm2 = __fsub_rn(0.0f, __fmul_rn(x, x));
    float e = reduced_range_exp(xm2);
    float r = __fsub_rn(1.0f, __fmul_rn(at, e));
    return neg ? __fsub_rn(0.0f, r) : r;
}

__device__ __forceinline__ float gelu_ref(float x) {
    float half_x = __fmul_rn(x, 0.5f);
    float y = __fmul_rn(x, 0.70710678118654752440f);
    y = __fadd_rn(erf_ref(y), 1.0f);
    return __fmul_rn(half_x, y);
}

__device__ __forceinline__ float tanh_ref(float x) {
    bool neg = x <= 0.0f;
    float ax = fabsf(x);
    float x2 = __fmul_rn(x, x);
    float ys = __fmaf_rn(1.5497927553951740264892578125e-2f, x2, -5.21197654306888580322265625e-2f);
    ys = __fmaf_rn(ys, x2, 0.13310669362545013427734375f);
    ys = __fmaf_rn(ys, x2, -0.33332359790802001953125f);
    ys = __fmaf_rn(ys, x2, 0.999999940395355224609375f);
    ys = __fmul_rn(ys, ax);
    float e = exp_ref(__fmul_rn(ax, 2.0f));
    float ym = __fdiv_rn(__fsub_rn(e, 1.0f), __fadd_rn(e, 1.0f));
    float y = (ax >= 9.02f) ? 1.0f : ym;
    if (ax <= 0.55f) y = ys;
    if (ax <= 0.0004f) y = ax;
    return neg ? __fsub_rn(0.0f, y) : y;
}

__device__ __forceinline__ float approx_gelu_ref(float x) {
    float half_x = __fmul_rn(x, 0.5f);
    float x3 = __fmul_rn(__fmul_rn(x, x), x);
    float y = __fmaf_rn(x3, 0.044715f, x);
    y = __fmul_rn(y, 0.7978845608028654f);
    y = tanh_ref(y);
    y = __fadd_rn(y, 1.0f);
    return __fmul_rn(half_x, y);
}

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case 1: return v > 0.0f ? v : 0.0f;
        case 2: return gelu_ref(v);
        case 3: return approx_gelu_ref(v);
        default: return v;
    }
}

}  // namespace rtb
