// Device helpers shared by the row kernels (rowops.cu) and the fused skinny-M kernels (skinny.cu): the exact scalar
// recipes of DynamicQuantizeLinear (src/ops/quantize.rs:352-434, rten-vecmath/src/quantize.rs:38-77) and the fold step
// of Sum / SumSquareSub (rten-vecmath/src/sum.rs:22-35,111-130).
#pragma once
#include <cstdint>

namespace rtb {

__device__ __forceinline__ int float_to_ordered(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__device__ __forceinline__ void dql_params(const int* mm, float& scale, float& inv_scale, int& zp) {
    const float x_min = ordered_to_float(mm[0]), x_max = ordered_to_float(mm[1]);
    const float lo = fminf(x_min, 0.0f), hi = fmaxf(x_max, 0.0f);
    scale = __fdiv_rn(__fsub_rn(hi, lo), 255.0f);
    const float min_scaled = __fdiv_rn(lo, scale);
    float z = __fsub_rn(0.0f, min_scaled);
    z = fminf(fmaxf(z, 0.0f), 255.0f);  // clamp (NaN -> 0 after the cast below)
    z = rintf(z);                       // round_ties_even
    zp = (z != z) ? 0 : (int)z;
    inv_scale = __fdiv_rn(1.0f, scale);
}

__device__ __forceinline__ int rne_i32_x86(float v) {
    if (!(v >= -2147483648.0f && v < 2147483648.0f)) return (int)0x80000000;
    return __float2int_rn(v);
}
__device__ __forceinline__ uint8_t quant1(float x, float inv_scale, int zp) {
    // saturate_u8(round(x * inv_scale) + zp): the rounded value is clamped to [-256, 511] first -- anything outside
    // saturates the same way -- so that the sum stays in 32 bits (zp is in [0, 255])
    int r = rne_i32_x86(__fmul_rn(x, inv_scale));
    r = max(-256, min(511, r));
    return (uint8_t)max(0, min(255, r + zp));
}

template <bool SQSUB>
__device__ __forceinline__ float fold_step(float acc, float x, float off) {
    if (SQSUB) {
        const float d = __fsub_rn(x, off);
        return __fmaf_rn(d, d, acc);
    }
    return __fadd_rn(acc, x);
}

// Sum / SumSquareSub of one row held in registers as float4s, in the reference's fold_unroll<4> x 16-lane order (see
// layer_norm_vec_kernel in rowops.cu for the thread <-> chain mapping): thread (c = lane & 15, segment seg) of a row
// that spans 16 S lanes holds the float4s f = c + 16 (seg F + k), k < F.  Every lane of the row returns the total.
template <int S, bool SQSUB, int FMAX = 16>
__device__ __forceinline__ float ln_vec_fold(const float4 (&v)[FMAX], int F, float off, int c, int seg) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int sg = 0; sg < S; sg++) {
        if (sg > 0) {
            const float4 in = make_float4(__shfl_up_sync(0xffffffffu, acc.x, 16), __shfl_up_sync(0xffffffffu, acc.y, 16),
                                          __shfl_up_sync(0xffffffffu, acc.z, 16), __shfl_up_sync(0xffffffffu, acc.w, 16));
            if (seg == sg) acc = in;
        }
        if (seg == sg) {
#pragma unroll
            for (int k = 0; k < FMAX; k++) {
                if (k < F) {
                    acc.x = fold_step<SQSUB>(acc.x, v[k].x, off);
                    acc.y = fold_step<SQSUB>(acc.y, v[k].y, off);
                    acc.z = fold_step<SQSUB>(acc.z, v[k].z, off);
                    acc.w = fold_step<SQSUB>(acc.w, v[k].w, off);
                }
            }
        }
    }
    // u = c >> 2 selects the unrolled accumulator, l = 4 (c & 3) + j the lane: threads c, c + 4, c + 8, c + 12 -> thread c (< 4)
    float4 r = acc;
#pragma unroll
    for (int u = 1; u < 4; u++) {
        r.x = __fadd_rn(r.x, __shfl_down_sync(0xffffffffu, acc.x, 4 * u));
        r.y = __fadd_rn(r.y, __shfl_down_sync(0xffffffffu, acc.y, 4 * u));
        r.z = __fadd_rn(r.z, __shfl_down_sync(0xffffffffu, acc.z, 4 * u));
        r.w = __fadd_rn(r.w, __shfl_down_sync(0xffffffffu, acc.w, 4 * u));
    }
    float s = 0.0f;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const float in = __shfl_up_sync(0xffffffffu, s, 1);
        if (c == q) {
            if (q > 0) s = in;
            s = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s, r.x), r.y), r.z), r.w);
        }
    }
    // thread c = 3 of the row's LAST segment holds the total
    const int lane = threadIdx.x & 31;
    const int base = (lane / (16 * S)) * (16 * S);
    return __shfl_sync(0xffffffffu, s, base + (S - 1) * 16 + 3);
}

}  // namespace rtb
