// Skinny-M kernels for autoregressive decode: HBM-streaming vector-matrix products with the surrounding operators fused
// in, and single-query attention over a KV cache.  See skinny.h.
//
// These replace, for M <= 16 (int8) / M <= 32 (f32) rows, the reference's gemv path (rten-gemm/src/lib.rs:668-747,
// rten-gemm/src/kernels/simd_generic.rs:14-197,795-1129): the weight matrix is read exactly once from HBM with
// 128-bit loads, the activations sit in shared memory, and there is no tensor-core tile to pad M up to.
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <cstdint>
#include <cstdlib>

#include "math.cuh"
#include "rowmath.cuh"
#include "skinny.h"

namespace rtb {

namespace {

__device__ __forceinline__ int dp4a_us(unsigned a, unsigned b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp4a_uu(unsigned a, unsigned b, int c) {
    unsigned d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"((unsigned)c));
    return (int)d;
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Sum of v[i] over the 32 lanes for NV values with ~NV shuffles instead of 5 NV: at every step a lane hands half of
// its values to its partner and keeps (and accumulates) the other half.  On return lane `l` holds `nout` complete sums,
// v[0 .. nout), of the indices base .. base + nout - 1.  Integer or float; the float order is fixed (deterministic).
template <int NV, typename T>
__device__ __forceinline__ void reduce_scatter_warp(T (&v)[NV], int lane, int& base, int& nout) {
    base = 0;
    int n = NV;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        if (n > 1) {
            const int half = n >> 1;
            const bool upper = (lane & o) != 0;
#pragma unroll
            for (int i = 0; i < NV / 2; i++) {
                if (i < half) {
                    const T send = upper ? v[i] : v[i + half];
                    const T keep = upper ? v[i + half] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            }
            if (upper) base += half;
            n = half;
        } else {
            v[0] = v[0] + __shfl_xor_sync(0xffffffffu, v[0], o);
        }
    }
    nout = n;
}

}  // namespace

// =========================================================================================
// Fused [LayerNorm] -> DynamicQuantizeLinear -> int8 GEMV -> scale / bias / residual / activation
// =========================================================================================
struct QLinearParams {
    QLinearLaunch L;
    int tiles;  // column tiles of 8 * CPW columns
};

// One row of x, optionally layer-normalised, as float4s in the vector-LayerNorm mapping (32 lanes per row):
// thread (c = lane & 15, seg = lane >> 4) holds the float4s f = c + 16 (seg F + k), k < F = K / 128.
__device__ __forceinline__ void qlin_ln_row(const QLinearLaunch& L, int r, int lane, float4 (&v)[16]) {
    const int c = lane & 15, seg = lane >> 4;
    const int F = L.K >> 7;
    const float4* x4 = reinterpret_cast<const float4*>(L.x + (long long)r * L.xs);
#pragma unroll
    for (int k = 0; k < 16; k++)
        if (k < F) v[k] = x4[c + 16 * (seg * F + k)];
    const float mean = __fdiv_rn(ln_vec_fold<2, false>(v, F, 0.0f, c, seg), (float)L.K);
    const float var = __fdiv_rn(ln_vec_fold<2, true>(v, F, mean, c, seg), (float)L.K);
    const float rstd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, L.ln_eps)));
    const float4* g4 = reinterpret_cast<const float4*>(L.ln_gamma);
    const float4* b4 = reinterpret_cast<const float4*>(L.ln_beta);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k < F) {
            const int f = c + 16 * (seg * F + k);
            const float4 a = v[k];
            const float4 g = __ldg(g4 + f);
            if (!L.ln_beta) {  // (same arm as layer_norm_vec_kernel's mode 1)
                v[k] = make_float4(__fmul_rn(__fsub_rn(a.x, mean), __fmul_rn(g.x, rstd)), __fmul_rn(__fsub_rn(a.y, mean), __fmul_rn(g.y, rstd)),
                                   __fmul_rn(__fsub_rn(a.z, mean), __fmul_rn(g.z, rstd)), __fmul_rn(__fsub_rn(a.w, mean), __fmul_rn(g.w, rstd)));
            } else {  // (mode 2: beta + the scalar bias 0.0)
                const float4 b = __ldg(b4 + f);
                v[k] = make_float4(__fmaf_rn(__fsub_rn(a.x, mean), __fmul_rn(g.x, rstd), __fadd_rn(b.x, 0.0f)),
                                   __fmaf_rn(__fsub_rn(a.y, mean), __fmul_rn(g.y, rstd), __fadd_rn(b.y, 0.0f)),
                                   __fmaf_rn(__fsub_rn(a.z, mean), __fmul_rn(g.z, rstd), __fadd_rn(b.z, 0.0f)),
                                   __fmaf_rn(__fsub_rn(a.w, mean), __fmul_rn(g.w, rstd), __fadd_rn(b.w, 0.0f)));
            }
        }
    }
}

__device__ __forceinline__ uint32_t quant4(float4 a, float inv, int zp) {
    return (uint32_t)quant1(a.x, inv, zp) | ((uint32_t)quant1(a.y, inv, zp) << 8) | ((uint32_t)quant1(a.z, inv, zp) << 16) |
           ((uint32_t)quant1(a.w, inv, zp) << 24);
}

template <int MT, int CPW, bool WSIGNED>
__global__ void __launch_bounds__(256) qlinear_kernel(const QLinearParams p) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    const QLinearLaunch& L = p.L;
    const int K = L.K, M = L.M, N = L.N;
    uint8_t* aq = sm_raw;  // [MT][K]
    int* s_rowsum = reinterpret_cast<int*>(aq + (size_t)MT * K);
    float* s_lo = reinterpret_cast<float*>(s_rowsum + MT);
    float* s_hi = s_lo + 8;
    int* s_mm = reinterpret_cast<int*>(s_hi + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NT = 8 * CPW;
    const uint8_t* w = reinterpret_cast<const uint8_t*>(L.w);

    // the weights do not depend on the previous kernel: pull this CTA's first tile towards L2 while it still runs
    {
        const long long lines = ((long long)NT * K + 127) >> 7;
        const int n0 = blockIdx.x * NT;
        for (long long i = tid; i < lines; i += 256) {
            const long long byte = i << 7;
            const int n = n0 + (int)(byte / K);
            if (n < N) prefetch_l2(w + (long long)n * L.ldw + (byte % K));
        }
    }
    pdl_wait();
    pdl_launch_dependents();

    // ---- pass 1: range of the (normalised) input
    float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
    if (L.has_ln) {
        for (int r = warp; r < M; r += 8) {
            float4 v[16];
            qlin_ln_row(L, r, lane, v);
            const int F = K >> 7;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (k < F) {
                    lo = fminf(fminf(lo, v[k].x), fminf(v[k].y, fminf(v[k].z, v[k].w)));
                    hi = fmaxf(fmaxf(hi, v[k].x), fmaxf(v[k].y, fmaxf(v[k].z, v[k].w)));
                }
            }
        }
    } else {
        const int k4 = K >> 2;
        for (int i = tid; i < M * k4; i += 256) {
            const int r = i / k4, f = i - r * k4;
            const float4 a = reinterpret_cast<const float4*>(L.x + (long long)r * L.xs)[f];
            lo = fminf(fminf(lo, a.x), fminf(a.y, fminf(a.z, a.w)));
            hi = fmaxf(fmaxf(hi, a.x), fmaxf(a.y, fmaxf(a.z, a.w)));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) {
        s_lo[warp] = lo;
        s_hi[warp] = hi;
    }
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < 8; k++) {
            lo = fminf(lo, s_lo[k]);
            hi = fmaxf(hi, s_hi[k]);
        }
        s_mm[0] = float_to_ordered(lo);
        s_mm[1] = float_to_ordered(hi);
    }
    __syncthreads();
    float x_scale, inv;
    int zp;
    dql_params(s_mm, x_scale, inv, zp);

    // ---- pass 2: quantise into shared memory (rows >= M are zero)
    uint32_t* aq32 = reinterpret_cast<uint32_t*>(aq);
    const int k4 = K >> 2;
    if (L.has_ln) {
        for (int r = warp; r < M; r += 8) {
            float4 v[16];
            qlin_ln_row(L, r, lane, v);
            const int c = lane & 15, seg = lane >> 4, F = K >> 7;
#pragma unroll
            for (int k = 0; k < 16; k++)
                if (k < F) aq32[r * k4 + c + 16 * (seg * F + k)] = quant4(v[k], inv, zp);
        }
    } else {
        for (int i = tid; i < M * k4; i += 256) {
            const int r = i / k4, f = i - r * k4;
            aq32[i] = quant4(reinterpret_cast<const float4*>(L.x + (long long)r * L.xs)[f], inv, zp);
        }
    }
    for (int i = M * k4 + tid; i < MT * k4; i += 256) aq32[i] = 0u;
    __syncthreads();
    if (L.zb) {  // row sums of the quantised activations for the weight-zero-point term
        for (int r = warp; r < MT; r += 8) {
            int s = 0;
            for (int i = lane; i < k4; i += 32) s = dp4a_uu(aq32[r * k4 + i], 0x01010101u, s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) s_rowsum[r] = s;
        }
        __syncthreads();
    }

    // ---- GEMV: warp `warp` of tile t owns columns t * NT + warp * CPW .. + CPW - 1; lanes split K in 16-byte chunks
    const uint4* aq4 = reinterpret_cast<const uint4*>(aq);
    const int KC = K >> 4;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const int n0 = tile * NT + warp * CPW;
        int acc[MT * CPW];
#pragma unroll
        for (int i = 0; i < MT * CPW; i++) acc[i] = 0;
        if (n0 < N) {
#pragma unroll 2
            for (int c = lane; c < KC; c += 32) {
                uint4 wv[CPW];
#pragma unroll
                for (int j = 0; j < CPW; j++) {
                    const int n = n0 + j < N ? n0 + j : N - 1;  // (clamped: the duplicate column is never stored)
                    wv[j] = __ldg(reinterpret_cast<const uint4*>(w + (long long)n * L.ldw) + c);
                }
#pragma unroll
                for (int m = 0; m < MT; m++) {
                    const uint4 av = aq4[m * KC + c];
#pragma unroll
                    for (int j = 0; j < CPW; j++) {
                        int a = acc[m * CPW + j];
                        if (WSIGNED) {
                            a = dp4a_us(av.x, wv[j].x, a);
                            a = dp4a_us(av.y, wv[j].y, a);
                            a = dp4a_us(av.z, wv[j].z, a);
                            a = dp4a_us(av.w, wv[j].w, a);
                        } else {
                            a = dp4a_uu(av.x, wv[j].x, a);
                            a = dp4a_uu(av.y, wv[j].y, a);
                            a = dp4a_uu(av.z, wv[j].z, a);
                            a = dp4a_uu(av.w, wv[j].w, a);
                        }
                        acc[m * CPW + j] = a;
                    }
                }
            }
        }
        int base, nout;
        reduce_scatter_warp<MT * CPW>(acc, lane, base, nout);
        // ---- epilogue: lane holds the exact i32 dot products of (m, j) = divmod(base + i, CPW)
#pragma unroll
        for (int i = 0; i < (MT * CPW + 31) / 32; i++) {
            if (i < nout) {
                const int idx = base + i;
                const int m = idx / CPW, n = n0 + idx % CPW;
                if (m < M && n < N) {
                    // C = acc - za*colsum[n] - zb[n]*(rowsum[m] - K*za), wrapping 32-bit (rten-gemm/src/kernels/simd_generic.rs:676-746)
                    unsigned cval = (unsigned)acc[i] - (unsigned)zp * (unsigned)__ldg(L.colsum + n);
                    if (L.zb) {
                        const unsigned zbv = (unsigned)__ldg(L.zb + (L.zb_len == 1 ? 0 : n));
                        cval -= zbv * ((unsigned)s_rowsum[m] - (unsigned)K * (unsigned)zp);
                    }
                    // Mul(x_scale, w_scale), cast * scale, Add(bias), Add(residual), activation: separate exactly rounded ops
                    const float sc = __fmul_rn(x_scale, __ldg(L.w_scale + (L.w_scale_len == 1 ? 0 : n)));
                    float xv = __fmul_rn(__int2float_rn((int)cval), sc);
                    if (L.bias) xv = __fadd_rn(xv, __ldg(L.bias + n));
                    if (L.residual) xv = __fadd_rn(xv, L.residual[(long long)m * L.rs + n]);
                    L.out[(long long)m * L.os + n] = apply_act(xv, L.act);
                }
            }
        }
    }
}

bool qlinear_supported(const QLinearLaunch& L) {
    if (getenv("RTEN_B200_NO_SKINNY")) return false;
    if (L.M < 1 || L.M > 16 || L.N < 1 || L.K < 16 || (L.K & 15)) return false;
    if ((size_t)16 * L.K + 256 > 200 * 1024) return false;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!al16(L.x) || (L.xs & 3) || !al16(L.w) || (L.ldw & 15)) return false;
    if (L.has_ln && ((L.K & 127) || (L.K >> 7) > 16 || !L.ln_gamma || !al16(L.ln_gamma) || !al16(L.ln_beta))) return false;
    if (!L.colsum || !L.w_scale) return false;
    return true;
}

rten_status launch_qlinear(rten_ctx* ctx, const QLinearLaunch& L) {
    QLinearParams p;
    p.L = L;
    // columns per warp: more columns in flight per warp for wide outputs (bytes in flight per SM), fewer for narrow
    // ones so that the tiles still cover the SMs
    const int mt = L.M <= 8 ? 8 : 16;
    int cpw = L.N >= 16384 ? 8 : (L.N >= 2048 ? 2 : 1);
    if (mt == 16 && cpw == 8) cpw = 4;
    const int nt = 8 * cpw;
    p.tiles = (L.N + nt - 1) / nt;
    const int grid = std::min(p.tiles, 2 * ctx->num_sms);
    const size_t smem = (size_t)mt * L.K + mt * sizeof(int) + 16 * sizeof(float) + 16;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = launch_stream(ctx);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = getenv("RTEN_B200_NO_PDL") ? 0 : 1;
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        return cudaLaunchKernelEx(&cfg, kern, p);
    };
    cudaError_t e;
    const int key = (mt == 16 ? 100 : 0) + cpw * 2 + (L.w_signed ? 1 : 0);
    switch (key) {
        case 2: e = go(qlinear_kernel<8, 1, false>); break;
        case 3: e = go(qlinear_kernel<8, 1, true>); break;
        case 4: e = go(qlinear_kernel<8, 2, false>); break;
        case 5: e = go(qlinear_kernel<8, 2, true>); break;
        case 16: e = go(qlinear_kernel<8, 8, false>); break;
        case 17: e = go(qlinear_kernel<8, 8, true>); break;
        case 102: e = go(qlinear_kernel<16, 1, false>); break;
        case 103: e = go(qlinear_kernel<16, 1, true>); break;
        case 104: e = go(qlinear_kernel<16, 2, false>); break;
        case 105: e = go(qlinear_kernel<16, 2, true>); break;
        case 108: e = go(qlinear_kernel<16, 4, false>); break;
        default: e = go(qlinear_kernel<16, 4, true>); break;
    }
    if (e != cudaSuccess) return fail_cuda(ctx, e, "qlinear launch");
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(ctx, e, "qlinear launch");
    count_launch(ctx);
    return RTEN_OK;
}

// =========================================================================================
// f32 skinny GEMM (exact FMA arithmetic): A staged through shared memory in K chunks, B streamed once
// =========================================================================================
struct SkinnyF32Params {
    SkinnyF32Launch L;
    int tiles, kc;  // column tiles of 8 * CPW columns; K chunk (floats) held in shared memory
};

template <int MT, int CPW>
__global__ void __launch_bounds__(256) skinny_f32_kernel(const SkinnyF32Params p) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    const SkinnyF32Launch& L = p.L;
    float4* as4 = reinterpret_cast<float4*>(sm_raw);  // [MT][kc / 4]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NT = 8 * CPW;
    const int K = L.K, M = L.M, N = L.N;
    pdl_wait();
    pdl_launch_dependents();
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const int n0 = tile * NT + warp * CPW;
        float acc[MT * CPW];
#pragma unroll
        for (int i = 0; i < MT * CPW; i++) acc[i] = 0.0f;
        for (int k0 = 0; k0 < K; k0 += p.kc) {
            const int kn = min(p.kc, K - k0);  // multiple of 4
            const int q4 = kn >> 2, ld4 = p.kc >> 2;
            __syncthreads();  // the previous chunk has been consumed
            for (int i = tid; i < MT * q4; i += 256) {
                const int r = i / q4, f = i - r * q4;
                as4[r * ld4 + f] = r < M ? reinterpret_cast<const float4*>(L.a + (long long)r * L.as + k0)[f] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncthreads();
            if (n0 < N) {
#pragma unroll 2
                for (int c = lane; c < q4; c += 32) {
                    float4 wv[CPW];
#pragma unroll
                    for (int j = 0; j < CPW; j++) {
                        const int n = n0 + j < N ? n0 + j : N - 1;
                        wv[j] = __ldg(reinterpret_cast<const float4*>(L.b + (long long)n * L.bs + k0) + c);
                    }
#pragma unroll
                    for (int m = 0; m < MT; m++) {
                        const float4 av = as4[m * ld4 + c];
#pragma unroll
                        for (int j = 0; j < CPW; j++) {
                            float a = acc[m * CPW + j];
                            a = __fmaf_rn(av.x, wv[j].x, a);
                            a = __fmaf_rn(av.y, wv[j].y, a);
                            a = __fmaf_rn(av.z, wv[j].z, a);
                            a = __fmaf_rn(av.w, wv[j].w, a);
                            acc[m * CPW + j] = a;
                        }
                    }
                }
            }
        }
        int base, nout;
        reduce_scatter_warp<MT * CPW>(acc, lane, base, nout);
#pragma unroll
        for (int i = 0; i < (MT * CPW + 31) / 32; i++) {
            if (i < nout) {
                const int idx = base + i;
                const int m = idx / CPW, n = n0 + idx % CPW;
                if (m < M && n < N) {
                    // same epilogue arithmetic as the tensor-core kernel: act(alpha * acc + r_scale * R + bias)
                    float xv = acc[i] * L.alpha;
                    if (L.residual) xv = fmaf(L.r_scale, L.residual[(long long)m * L.rs + n], xv);
                    if (L.bias) xv = xv + __ldg(L.bias + n);
                    L.out[(long long)m * L.os + n] = apply_act(xv, L.act);
                }
            }
        }
    }
}

bool skinny_f32_supported(const SkinnyF32Launch& L) {
    if (getenv("RTEN_B200_NO_SKINNY")) return false;
    if (L.M < 1 || L.M > 32 || L.N < 1 || L.K < 4 || (L.K & 3)) return false;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    return al16(L.a) && al16(L.b) && !(L.as & 3) && !(L.bs & 3);
}

rten_status launch_skinny_f32(rten_ctx* ctx, const SkinnyF32Launch& L) {
    SkinnyF32Params p;
    p.L = L;
    const int mt = L.M <= 8 ? 8 : (L.M <= 16 ? 16 : 32);
    const int cpw = mt == 32 ? 1 : (L.N >= 4096 ? 2 : 1);
    const int nt = 8 * cpw;
    p.tiles = (L.N + nt - 1) / nt;
    p.kc = std::min((L.K + 3) / 4 * 4, mt == 32 ? 1024 : 2048);
    const int grid = std::min(p.tiles, 2 * ctx->num_sms);
    const size_t smem = (size_t)mt * p.kc * sizeof(float);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = launch_stream(ctx);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = getenv("RTEN_B200_NO_PDL") ? 0 : 1;
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) return e;
        return cudaLaunchKernelEx(&cfg, kern, p);
    };
    cudaError_t e;
    if (mt == 8)
        e = cpw == 2 ? go(skinny_f32_kernel<8, 2>) : go(skinny_f32_kernel<8, 1>);
    else if (mt == 16)
        e = cpw == 2 ? go(skinny_f32_kernel<16, 2>) : go(skinny_f32_kernel<16, 1>);
    else
        e = go(skinny_f32_kernel<32, 1>);
    if (e != cudaSuccess) return fail_cuda(ctx, e, "skinny f32 launch");
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(ctx, e, "skinny f32 launch");
    count_launch(ctx);
    return RTEN_OK;
}

// =========================================================================================
// Single-query attention over a KV cache (flash-decoding split over the cached sequence)
// =========================================================================================
struct AttnDecodeParams {
    AttnDecodeLaunch L;
    int nsplit;
    float* ws;  // [B * q_heads][nsplit][2 + dh]  partial (max, sum, unnormalised output)
    int* cnt;   // [B * q_heads] arrival counters (zero between launches)
};

constexpr int ATTN_MAX_CHUNK = 4096;  // positions of one split held in shared memory

template <int DH>
__global__ void __launch_bounds__(256) attn_decode_kernel(const AttnDecodeParams p) {
    __shared__ float s_p[ATTN_MAX_CHUNK];
    __shared__ float s_q[DH];
    __shared__ float s_red[8];
    __shared__ float s_o[8][DH];
    __shared__ float s_bcast[2];
    __shared__ int s_last;
    const AttnDecodeLaunch& L = p.L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bh = blockIdx.x / p.nsplit, split = blockIdx.x - bh * p.nsplit;
    const int b = bh / L.q_heads, h = bh - b * L.q_heads;
    const int hk = h / (L.q_heads / L.kv_heads);
    pdl_wait();
    pdl_launch_dependents();
    int len = L.len ? L.len[b] : L.kv_cap;
    len = max(0, min(len, L.kv_cap));
    const int per = (len + p.nsplit - 1) / p.nsplit;
    const int l0 = min(len, split * per), l1 = min(len, l0 + per);
    const int nl = l1 - l0;
    float* kc = L.k + (long long)b * L.k_b + (long long)hk * L.k_h;
    float* vc = L.v + (long long)b * L.v_b + (long long)hk * L.v_h;
    if (tid < DH) s_q[tid] = L.q[(long long)b * L.q_b + (long long)h * L.q_h + tid];
    // fused cache append: the split that owns position len - 1 writes the new key / value there first (one CTA per
    // kv head does it: the query heads of a group share the cache row)
    const bool appends = L.k_new && len > 0 && l1 == len && nl > 0 && (h % (L.q_heads / L.kv_heads)) == 0;
    if (appends && tid < DH) {
        kc[(long long)(len - 1) * L.k_l + tid] = L.k_new[(long long)b * L.kn_b + (long long)hk * L.kn_h + tid];
        vc[(long long)(len - 1) * L.v_l + (long long)tid * L.v_d] = L.v_new[(long long)b * L.vn_b + (long long)hk * L.vn_h + tid];
    }
    __syncthreads();
    // ---- scores: 8 lanes per cached position (DH / 8 floats each), 4 positions per warp and iteration
    constexpr int PER_LANE = DH / 8;  // 8 (dh 64) or 16 (dh 128) floats
    const int sub = lane >> 3, l8 = lane & 7;
    float qreg[PER_LANE];
#pragma unroll
    for (int i = 0; i < PER_LANE; i++) qreg[i] = s_q[l8 * PER_LANE + i];
    const float* mrow = L.mask ? L.mask + (long long)b * L.m_b + (long long)h * L.m_h : nullptr;
    const bool new_in_regs = L.k_new != nullptr;  // (other query heads of the group may race with the append: read k_new)
    float mx = -FLT_MAX;
    for (int i0 = warp * 4; i0 < nl; i0 += 32) {
        const int i = i0 + sub;
        float s = 0.0f;
        if (i < nl) {
            const int l = l0 + i;
            const float* kr = (new_in_regs && l == len - 1) ? L.k_new + (long long)b * L.kn_b + (long long)hk * L.kn_h
                                                            : kc + (long long)l * L.k_l;
            const float4* k4 = reinterpret_cast<const float4*>(kr + l8 * PER_LANE);
#pragma unroll
            for (int j = 0; j < PER_LANE / 4; j++) {
                const float4 kv = k4[j];
                s = fmaf(qreg[4 * j], kv.x, s);
                s = fmaf(qreg[4 * j + 1], kv.y, s);
                s = fmaf(qreg[4 * j + 2], kv.z, s);
                s = fmaf(qreg[4 * j + 3], kv.w, s);
            }
        }
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (i < nl) {
            s *= L.scale;
            if (mrow) s += mrow[(long long)(l0 + i) * L.m_l];
            if (l8 == 0) s_p[i] = s;
            mx = fmaxf(mx, s);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    if (tid == 0) {
        float m = s_red[0];
        for (int k = 1; k < 8; k++) m = fmaxf(m, s_red[k]);
        s_bcast[0] = m;
    }
    __syncthreads();
    mx = s_bcast[0];
    // ---- exponentials (the reference's polynomial, rten-vecmath/src/exp.rs:140-191) and their sum
    float sum = 0.0f;
    for (int i = tid; i < nl; i += 256) {
        const float e = reduced_range_exp(s_p[i] - mx);
        s_p[i] = e;
        sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();  // (s_red is reused; every thread has read s_bcast[0])
    if (lane == 0) s_red[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        float t = 0.0f;
        for (int k = 0; k < 8; k++) t += s_red[k];
        s_bcast[1] = t;
    }
    __syncthreads();
    sum = s_bcast[1];
    // ---- unnormalised output o[d] = sum_l p[l] V[l, d]
    const float* vnew = L.v_new ? L.v_new + (long long)b * L.vn_b + (long long)hk * L.vn_h : nullptr;
    float o_mine = 0.0f;  // thread d < DH ends up with o[d]
    if (L.v_l == 1) {
        // transposed cache [.., dh, cap]: warp w owns rows d = w, w + 8, ...; lanes stride the cached positions
        for (int d = warp; d < DH; d += 8) {
            const float* vr = vc + (long long)d * L.v_d + l0;
            float a = 0.0f;
            for (int i = lane; i < nl; i += 32) {
                const float vv = (vnew && l0 + i == len - 1) ? vnew[d] : vr[i];
                a = fmaf(s_p[i], vv, a);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) s_o[0][d] = a;
        }
        __syncthreads();
        if (tid < DH) o_mine = s_o[0][tid];
    } else {
        // natural cache [.., cap, dh]: warp w takes positions w, w + 8, ...; lanes own dh / 32 consecutive channels
        constexpr int CH = DH / 32;
        float a[CH];
#pragma unroll
        for (int j = 0; j < CH; j++) a[j] = 0.0f;
        for (int i = warp; i < nl; i += 8) {
            const float pw = s_p[i];
            const float* vr = (vnew && l0 + i == len - 1) ? vnew : vc + (long long)(l0 + i) * L.v_l;
#pragma unroll
            for (int j = 0; j < CH; j++) a[j] = fmaf(pw, vr[(long long)(lane * CH + j) * L.v_d], a[j]);
        }
#pragma unroll
        for (int j = 0; j < CH; j++) s_o[warp][lane * CH + j] = a[j];
        __syncthreads();
        if (tid < DH) {
            float t = 0.0f;
            for (int k = 0; k < 8; k++) t += s_o[k][tid];
            o_mine = t;
        }
    }
    float* outp = L.out + (long long)b * L.o_b + (long long)h * L.o_h;
    if (p.nsplit == 1) {
        if (tid < DH) {
            float r = o_mine / sum;
            if (r != r) r = 0.0f;  // fully masked row -> zeros (sdpa_head flushes NaNs)
            outp[tid] = r;
        }
        return;
    }
    // ---- merge the splits: the last CTA of (b, h) to arrive combines the partial (max, sum, output) triples
    float* wsp = p.ws + ((long long)bh * p.nsplit + split) * (2 + DH);
    if (tid == 0) {
        wsp[0] = mx;
        wsp[1] = sum;
    }
    if (tid < DH) wsp[2 + tid] = o_mine;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int old = atomicAdd(p.cnt + bh, 1);
        s_last = old == p.nsplit - 1;
        if (s_last) p.cnt[bh] = 0;
        __threadfence();
    }
    __syncthreads();
    if (!s_last) return;
    if (tid < DH) {
        const float* w0 = p.ws + (long long)bh * p.nsplit * (2 + DH);
        float m = -FLT_MAX;
        for (int s2 = 0; s2 < p.nsplit; s2++) m = fmaxf(m, __ldcg(w0 + s2 * (2 + DH)));
        float num = 0.0f, den = 0.0f;
        for (int s2 = 0; s2 < p.nsplit; s2++) {
            const float* ww = w0 + s2 * (2 + DH);
            const float sc = reduced_range_exp(__ldcg(ww) - m);
            den = fmaf(__ldcg(ww + 1), sc, den);
            num = fmaf(__ldcg(ww + 2 + tid), sc, num);
        }
        float r = num / den;
        if (r != r) r = 0.0f;
        outp[tid] = r;
    }
}

bool attn_decode_supported(const AttnDecodeLaunch& L) {
    if (getenv("RTEN_B200_NO_SKINNY")) return false;
    if (L.dh != 64 && L.dh != 128) return false;
    if (L.B < 1 || L.q_heads < 1 || L.kv_heads < 1 || L.q_heads % L.kv_heads) return false;
    if (L.kv_cap < 1 || L.kv_cap > 8 * ATTN_MAX_CHUNK) return false;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!al16(L.k) || (L.k_b & 3) || (L.k_h & 3) || (L.k_l & 3)) return false;
    if (L.k_new && (!al16(L.k_new) || (L.kn_b & 3) || (L.kn_h & 3) || !L.v_new)) return false;
    if (L.v_l != 1 && L.v_d != 1) return false;
    return true;
}

rten_status launch_attn_decode(rten_ctx* ctx, const AttnDecodeLaunch& L) {
    AttnDecodeParams p;
    p.L = L;
    const int bh = L.B * L.q_heads;
    // enough CTAs to cover the SMs about twice, every split at least 64 positions and at most ATTN_MAX_CHUNK
    int ns = std::max(1, std::min(8, (2 * ctx->num_sms + bh - 1) / bh));
    ns = std::min(ns, std::max(1, L.kv_cap / 64));
    ns = std::max(ns, (L.kv_cap + ATTN_MAX_CHUNK - 1) / ATTN_MAX_CHUNK);
    p.nsplit = ns;
    p.ws = nullptr;
    p.cnt = nullptr;
    if (ns > 1) {
        if (!ctx->attn_cnt || ctx->attn_cnt_len < bh) {
            if (ctx->attn_cnt) cudaFree(ctx->attn_cnt);
            const int cap = std::max(bh, 1024);
            cudaError_t ce = cudaMalloc(&ctx->attn_cnt, (size_t)cap * sizeof(int));
            if (ce != cudaSuccess) return fail_cuda(ctx, ce, "attention counters");
            ce = cudaMemset(ctx->attn_cnt, 0, (size_t)cap * sizeof(int));
            if (ce != cudaSuccess) return fail_cuda(ctx, ce, "attention counters");
            ctx->attn_cnt_len = cap;
        }
        void* ws = nullptr;
        RTB_TRY(temp_alloc(ctx, (size_t)bh * ns * (2 + L.dh) * sizeof(float), &ws));
        p.ws = reinterpret_cast<float*>(ws);
        p.cnt = reinterpret_cast<int*>(ctx->attn_cnt);
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(bh * ns);
    cfg.blockDim = dim3(256);
    cfg.stream = launch_stream(ctx);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = getenv("RTEN_B200_NO_PDL") ? 0 : 1;
    cudaError_t e = L.dh == 64 ? cudaLaunchKernelEx(&cfg, attn_decode_kernel<64>, p) : cudaLaunchKernelEx(&cfg, attn_decode_kernel<128>, p);
    if (e != cudaSuccess) return fail_cuda(ctx, e, "attention launch");
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(ctx, e, "attention launch");
    count_launch(ctx);
    return RTEN_OK;
}

}  // namespace rtb
