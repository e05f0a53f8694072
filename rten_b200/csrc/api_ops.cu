// Operator entry points of the C ABI: shape/argument validation with the reference's error strings,
// operand normalisation (K-major, TMA-addressable), kernel dispatch.  Mirrors, per function, the
// reference operator named in include/rten_b200.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "api_util.h"
#include "rowops.h"
#include "skinny.h"
#include "umma_gemm.h"

using namespace rtb;

namespace {

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

rten_status check_ctx(rten_ctx* ctx) { return ctx ? RTEN_OK : RTEN_ERR_INVALID_VALUE; }

// ---------------------------------------------------------------------------------------
// K-major 2-level operand: rows x K with element strides.  Packs into an aligned workspace when
// TMA cannot address the original (k stride != 1, misaligned base / pitch).
// ---------------------------------------------------------------------------------------
struct Mat {
    const void* base;
    int64_t rows, K;
    int64_t rs, ks;       // element strides
    int64_t z0 = 1, z1 = 1;  // batch dims (z0 inner)
    int64_t zs0 = 0, zs1 = 0;
};

rten_status to_kmajor(rten_ctx* ctx, int esize, const Mat& m, OperandDesc* od) {
    OperandDesc d;
    d.base = m.base;
    d.dims[0] = m.K;
    d.dims[1] = m.rows;
    // broadcast batch dims (stride 0) become size-1 dims: the kernel then always passes coordinate 0
    d.dims[2] = (m.z0 > 1 && m.zs0 != 0) ? m.z0 : 1;
    d.dims[3] = (m.z1 > 1 && m.zs1 != 0) ? m.z1 : 1;
    d.strides[0] = m.ks;
    d.strides[1] = m.rs;
    d.strides[2] = d.dims[2] > 1 ? m.zs0 : 0;
    d.strides[3] = d.dims[3] > 1 ? m.zs1 : 0;
    if (m.K == 1) d.strides[0] = 1;  // a single k element is trivially contiguous
    if (tma_compatible(d, esize, 4)) {
        *od = d;
        return RTEN_OK;
    }
    // pack: [z1', z0', rows, Kpad]; broadcast batch dims (stride 0) are NOT expanded
    const int64_t kpad = round_up(m.K, 16 / esize);
    const int64_t e0 = (m.z0 > 1 && m.zs0 == 0) ? 1 : m.z0;
    const int64_t e1 = (m.z1 > 1 && m.zs1 == 0) ? 1 : m.z1;
    void* buf = nullptr;
    RTB_TRY(temp_alloc(ctx, (size_t)(e1 * e0 * m.rows * kpad) * esize, &buf));
    long long shape[4] = {e1, e0, m.rows, m.K};
    long long ss[4] = {m.zs1, m.zs0, m.rs, m.ks};
    long long ds[4] = {e0 * m.rows * kpad, m.rows * kpad, kpad, 1};
    RTB_TRY(launch_nd_copy(ctx, esize, m.base, buf, 4, shape, ss, ds));
    d.base = buf;
    d.dims[2] = e0;
    d.dims[3] = e1;
    d.strides[0] = 1;
    d.strides[1] = kpad;
    d.strides[2] = e0 > 1 ? m.rows * kpad : 0;
    d.strides[3] = e1 > 1 ? e0 * m.rows * kpad : 0;
    *od = d;
    return RTEN_OK;
}

// Collapse broadcast prefix dims of a matmul into at most 2 batch dims (z0 inner, z1 outer).  Dims are merged
// only when A, B and the output all advance uniformly across them.
struct BatchDims {
    int64_t z0 = 1, z1 = 1;
    int64_t a0 = 0, a1 = 0, b0 = 0, b1 = 0, o0 = 0, o1 = 0;
    bool ok = true;
};

BatchDims collapse_batch(const std::vector<int64_t>& size, const std::vector<int64_t>& as, const std::vector<int64_t>& bs,
                         const std::vector<int64_t>& os) {
    std::vector<int64_t> s, a, b, o;
    for (size_t i = 0; i < size.size(); i++) {
        if (size[i] == 1) continue;
        if (!s.empty() && a.back() == as[i] * size[i] && b.back() == bs[i] * size[i] && o.back() == os[i] * size[i]) {
            s.back() *= size[i];
            a.back() = as[i];
            b.back() = bs[i];
            o.back() = os[i];
            continue;
        }
        s.push_back(size[i]);
        a.push_back(as[i]);
        b.push_back(bs[i]);
        o.push_back(os[i]);
    }
    BatchDims r;
    if (s.size() > 2) {
        r.ok = false;
        return r;
    }
    if (s.size() == 1) {
        r.z0 = s[0];
        r.a0 = a[0];
        r.b0 = b[0];
        r.o0 = o[0];
    } else if (s.size() == 2) {
        r.z1 = s[0];
        r.a1 = a[0];
        r.b1 = b[0];
        r.o1 = o[0];
        r.z0 = s[1];
        r.a0 = a[1];
        r.b0 = b[1];
        r.o0 = o[1];
    }
    return r;
}

struct MatMulArgs {
    int kind;  // 0 f32, 1 int8
    const rten_tensor* a;
    const rten_tensor* b;
    const rten_packed* pb;
    EpilogueDesc epi;  // d / strides filled by matmul_core
    int out_dtype;
    // int8 extras
    const rten_tensor* a_zp = nullptr;
    const rten_tensor* b_zp = nullptr;
    // residual (same shape as out) for matmul_ex
    const rten_tensor* residual = nullptr;
};

// numpy-matmul shape logic of src/ops/matmul.rs:208-385 + kernel dispatch
rten_status matmul_core(OpScope& sc, MatMulArgs& A, rten_tensor* out) {
    rten_ctx* ctx = sc.ctx;
    rten_tensor a = *A.a, b = *A.b;
    if (a.ndim < 1 || b.ndim < 1) return fail(ctx, RTEN_ERR_INVALID_VALUE, "Inputs must have >= 1 dimensions");
    const bool a_vec = a.ndim == 1, b_vec = b.ndim == 1;
    if (a_vec) {  // [K] -> [1, K]
        a.ndim = 2;
        a.shape[1] = a.shape[0];
        a.strides[1] = a.strides[0];
        a.shape[0] = 1;
        a.strides[0] = 0;
    }
    if (b_vec) {  // [K] -> [K, 1]
        b.ndim = 2;
        b.shape[1] = 1;
        b.strides[1] = 0;
    }
    const int64_t M = a.shape[a.ndim - 2], K = a.shape[a.ndim - 1];
    const int64_t Kb = b.shape[b.ndim - 2], N = b.shape[b.ndim - 1];
    if (K != Kb)
        return fail(ctx, RTEN_ERR_INCOMPATIBLE_SHAPES, "Columns of first matrix does not match rows of second matrix");
    // broadcast prefixes
    const int pa = a.ndim - 2, pb = b.ndim - 2, pn = std::max(pa, pb);
    std::vector<int64_t> psize(pn), pas(pn), pbs(pn);
    for (int i = 0; i < pn; i++) {
        const int ia = i - (pn - pa), ib = i - (pn - pb);
        const int64_t sa = ia >= 0 ? a.shape[ia] : 1, sb = ib >= 0 ? b.shape[ib] : 1;
        if (sa != sb && sa != 1 && sb != 1) return fail(ctx, RTEN_ERR_INCOMPATIBLE_SHAPES, "Cannot broadcast shapes");
        psize[i] = std::max(sa, sb);
        if (sa == 0 || sb == 0) psize[i] = 0;
        pas[i] = (ia >= 0 && sa != 1) ? a.strides[ia] : 0;
        pbs[i] = (ib >= 0 && sb != 1) ? b.strides[ib] : 0;
    }
    // output shape
    int64_t oshape[RTEN_MAX_DIMS];
    int on = 0;
    for (int i = 0; i < pn; i++) oshape[on++] = psize[i];
    if (!a_vec) oshape[on++] = M;
    if (!b_vec) oshape[on++] = N;
    if (on > RTEN_MAX_DIMS) return fail(ctx, RTEN_ERR_INVALID_VALUE, "tensor rank out of range");
    rten_tensor ov;
    RTB_TRY(sc.out(out, A.out_dtype, on, oshape, &ov, nullptr));
    int64_t total = 1;
    for (int i = 0; i < on; i++) total *= oshape[i];
    if (total == 0) return RTEN_OK;
    // output strides of the prefix dims / row / col in the (vector-expanded) [prefix.., M, N] view
    std::vector<int64_t> pos(pn, 0);
    int64_t o_rs = 0, o_cs = 0;
    {
        int k = 0;
        for (int i = 0; i < pn; i++) pos[i] = ov.strides[k++];
        if (!a_vec) o_rs = ov.strides[k++];
        if (!b_vec) o_cs = ov.strides[k++];
    }
    rten_tensor dv = ov;
    bool copy_out = false;
    const int esize = A.kind == 0 ? 4 : 1;
    int64_t nbatch = 1;
    for (int i = 0; i < pn; i++) nbatch *= psize[i];
    int64_t nb_mats = 1;
    for (int i = 0; i < pb; i++) nb_mats *= b.shape[i];

    GemmLaunch L;
    L.kind = A.kind;
    L.a_signed = A.a->dtype == RTEN_I8;
    L.b_signed = (A.pb ? A.pb->dtype : A.b->dtype) == RTEN_I8;
    L.N = (int)N;
    L.K = (int)K;
    L.epi = A.epi;
    L.epi.d_is_i32 = A.out_dtype == RTEN_I32;

    if (K == 0) {
        // lib.rs:843-873: product term vanishes; out = bias (f32) / 0 (int).  Reuse Add machinery: fill.
        if (!is_contiguous(&ov)) {
            set_contiguous(&dv);
            void* t = nullptr;
            RTB_TRY(temp_alloc(ctx, (size_t)total * 4, &t));
            dv.data = t;
            copy_out = true;
        }
        RTB_CUDA(ctx, cudaMemsetAsync(dv.data, 0, (size_t)total * 4, rtb::launch_stream(ctx)));
        if (A.kind == 0 && L.epi.bias) {
            long long shp[2] = {total / N, N}, s0[2] = {N, 1}, sb[2] = {0, 1};
            RTB_TRY(launch_nd_add(ctx, (const float*)dv.data, L.epi.bias, (float*)dv.data, 2, shp, s0, sb, s0, 0));
        }
    } else {
        // B operand
        Mat mb;
        if (A.pb) {
            if (A.pb->kind != 0 || A.pb->K != K || A.pb->N != N)
                return fail(ctx, RTEN_ERR_INVALID_VALUE, "prepacked B does not match the matmul shape");
            mb.base = A.pb->data;
            mb.rows = N;
            mb.K = K;
            mb.rs = A.pb->ld;
            mb.ks = 1;
            nb_mats = 1;
        } else {
            mb.base = b.data;
            mb.rows = N;
            mb.K = K;
            mb.rs = b.strides[b.ndim - 1];
            mb.ks = b.strides[b.ndim - 2];
        }
        Mat ma;
        ma.base = a.data;
        ma.K = K;
        ma.ks = a.strides[a.ndim - 1];
        ma.rs = a.strides[a.ndim - 2];
        ma.rows = M;

        // Flatten [A.., M, K] x [K, N] into one [A*M, K] GEMM (matmul.rs:266-297) when A's rows are
        // uniformly strided; otherwise keep (up to two) batch dims.
        bool flat = false;
        BatchDims bd;
        const std::vector<int64_t> b_eff = A.pb ? std::vector<int64_t>(pn, 0) : pbs;
        for (int attempt = 0; attempt < 2; attempt++) {
            // attempt 0: write straight into the caller's (possibly strided) output; attempt 1: contiguous temp
            if (attempt == 1) {
                set_contiguous(&dv);
                void* t = nullptr;
                RTB_TRY(temp_alloc(ctx, (size_t)total * 4, &t));
                dv.data = t;
                copy_out = true;
                int k = 0;
                for (int i = 0; i < pn; i++) pos[i] = dv.strides[k++];
                if (!a_vec) o_rs = dv.strides[k++];
                if (!b_vec) o_cs = dv.strides[k++];
            }
            flat = false;
            if (nb_mats == 1) {
                std::vector<int64_t> sz = psize, as = pas, zs(pn, 0), os = pos;
                sz.push_back(M);
                as.push_back(ma.rs);
                zs.push_back(0);
                os.push_back(o_rs);
                BatchDims c = collapse_batch(sz, as, zs, os);
                if (c.ok && c.z1 == 1) {
                    flat = true;
                    ma.rows = c.z0;
                    if (c.z0 > 1) {
                        ma.rs = c.a0;
                        o_rs = c.o0;
                    }
                    break;
                }
            }
            bd = collapse_batch(psize, pas, b_eff, pos);
            if (bd.ok) break;
            if (attempt == 1)
                return fail(ctx, RTEN_ERR_UNSUPPORTED_VALUE, "matmul batch dims do not collapse to 2 strided dims");
        }
        L.epi.d = dv.data;
        L.epi.s_row = o_rs;
        L.epi.s_col = b_vec ? 1 : o_cs;
        if (flat) {
            L.M = (int)ma.rows;
            L.z0 = L.z1 = 1;
        } else {
            L.M = (int)M;
            L.z0 = (int)bd.z0;
            L.z1 = (int)bd.z1;
            ma.z0 = bd.z0;
            ma.z1 = bd.z1;
            ma.zs0 = bd.a0;
            ma.zs1 = bd.a1;
            mb.z0 = bd.z0;
            mb.z1 = bd.z1;
            mb.zs0 = bd.b0;
            mb.zs1 = bd.b1;
            L.epi.s_z0 = bd.o0;
            L.epi.s_z1 = bd.o1;
        }
        if (A.kind == 1 && !flat && (A.a_zp || A.b_zp) && nbatch > 1)
            return fail(ctx, RTEN_ERR_UNSUPPORTED_VALUE, "MatMulInteger with a batched RHS and zero points is not supported");
        RTB_TRY(to_kmajor(ctx, esize, ma, &L.a));
        if (mb.z0 > 1 && mb.zs0 == 0) {}  // broadcast handled by stride 0
        RTB_TRY(to_kmajor(ctx, esize, mb, &L.b));

        // residual (same shape as out, any strides) -> only contiguous or row-strided supported directly
        if (A.residual) {
            rten_tensor rv;
            RTB_TRY(sc.in(A.residual, &rv));
            if (rv.ndim != on) return fail(ctx, RTEN_ERR_INCOMPATIBLE_SHAPES, "residual shape does not match output");
            for (int i = 0; i < on; i++)
                if (rv.shape[i] != oshape[i]) return fail(ctx, RTEN_ERR_INCOMPATIBLE_SHAPES, "residual shape does not match output");
            rten_tensor rc;
            RTB_TRY(sc.contiguous(&rv, &rc));
            L.epi.r = (const float*)rc.data;
            L.epi.r_scale = 1.0f;
            L.epi.r_col = 1;
            L.epi.r_row = N;
            L.epi.r_z0 = flat ? 0 : M * N;
            L.epi.r_z1 = flat ? 0 : bd.z0 * M * N;
        }

        // integer zero points
        if (A.kind == 1) {
            const int64_t rows_total = flat ? ma.rows : M;
            if (A.a_zp) {
                rten_tensor z;
                RTB_TRY(sc.in(A.a_zp, &z));
                const int len = z.ndim == 0 ? 1 : (int)z.shape[0];
                if (len == 1) {  // scalar (DynamicQuantizeLinear's zero point): read in place by the epilogue
                    L.epi.za8 = (const uint8_t*)z.data;
                    L.epi.za8_signed = z.dtype == RTEN_I8;
                } else {
                    int32_t* za = nullptr;
                    RTB_TRY(temp_alloc(ctx, (size_t)len * 4, (void**)&za));
                    RTB_TRY(launch_zp_to_i32(ctx, z.data, z.dtype == RTEN_I8, len, z.ndim == 0 ? 0 : z.strides[0], za));
                    L.epi.za = za;
                    L.epi.za_len = len;
                }
                if (A.pb && A.pb->colsum) {
                    L.epi.colsum = A.pb->colsum;
                } else {
                    int32_t* cs = nullptr;
                    RTB_TRY(temp_alloc(ctx, (size_t)N * 4, (void**)&cs));
                    RTB_TRY(launch_rowsum8(ctx, L.b.base, L.b_signed, N, (int)K, L.b.strides[1], cs));
                    L.epi.colsum = cs;
                }
            }
            if (A.b_zp) {
                rten_tensor z;
                RTB_TRY(sc.in(A.b_zp, &z));
                const int len = z.ndim == 0 ? 1 : (int)z.shape[0];
                int32_t* zb = nullptr;
                RTB_TRY(temp_alloc(ctx, (size_t)len * 4, (void**)&zb));
                RTB_TRY(launch_zp_to_i32(ctx, z.data, z.dtype == RTEN_I8, len, z.ndim == 0 ? 0 : z.strides[0], zb));
                L.epi.zb = zb;
                L.epi.zb_len = len;
                int32_t* rs = nullptr;
                RTB_TRY(temp_alloc(ctx, (size_t)rows_total * 4, (void**)&rs));
                RTB_TRY(launch_rowsum8(ctx, L.a.base, L.a_signed, rows_total, (int)K, L.a.strides[1], rs));
                L.epi.rowsum = rs;
            }
        }
        // M <= 32 f32 rows: the HBM-streaming skinny kernel (exact f32 FMA arithmetic) instead of a 128-row MMA tile
        // (rten-gemm's gemv path, rten-gemm/src/lib.rs:668-747)
        if (A.kind == 0 && L.z0 == 1 && L.z1 == 1 && L.M <= 32 && L.epi.s_col == 1 && L.epi.bias_kind != 2 &&
            (!L.epi.r || L.epi.r_col == 1) && !L.epi.range && L.a.strides[0] == 1 && L.b.strides[0] == 1) {
            SkinnyF32Launch S;
            S.a = (const float*)L.a.base;
            S.as = L.a.strides[1];
            S.b = (const float*)L.b.base;
            S.bs = L.b.strides[1];
            S.M = L.M;
            S.N = L.N;
            S.K = L.K;
            S.alpha = L.epi.alpha;
            S.bias = L.epi.bias_kind == 1 ? L.epi.bias : nullptr;
            S.residual = L.epi.r;
            S.rs = L.epi.r_row;
            S.r_scale = L.epi.r_scale;
            S.act = L.epi.act;
            S.out = (float*)L.epi.d;
            S.os = L.epi.s_row;
            if (skinny_f32_supported(S)) {
                RTB_TRY(launch_skinny_f32(ctx, S));
                goto launched;
            }
        }
        {
        rten_status st = launch_umma_gemm(ctx, L);
        if (st == RTEN_ERR_UNSUPPORTED_VALUE) return fail(ctx, st, "GEMM operands are not addressable by TMA after packing");
        RTB_TRY(st);
        }
    launched:;
    }
    if (copy_out) {
        long long shape[RTEN_MAX_DIMS], ss[RTEN_MAX_DIMS], ds[RTEN_MAX_DIMS];
        for (int i = 0; i < on; i++) {
            shape[i] = oshape[i];
            ss[i] = dv.strides[i];
            ds[i] = ov.strides[i];
        }
        RTB_TRY(launch_nd_copy(ctx, 4, dv.data, ov.data, on, shape, ss, ds));
    }
    return RTEN_OK;
}

// src/ops/matmul.rs:513-533 zero_point_to_vec validation
rten_status check_zero_point(rten_ctx* ctx, const rten_tensor* zp, int64_t expected, int want_dtype) {
    if (!zp) return RTEN_OK;
    if (zp->dtype != want_dtype) return fail(ctx, RTEN_ERR_CAST_FAILED, "zero point type does not match its tensor");
    if (zp->ndim == 0) return RTEN_OK;
    if (zp->ndim == 1) {
        if (zp->shape[0] != expected) return fail(ctx, RTEN_ERR_INVALID_VALUE, "Zero point has incorrect size");
        return RTEN_OK;
    }
    return fail(ctx, RTEN_ERR_UNSUPPORTED_VALUE, "Only scalar or vector zero points are supported");
}

// ---------------------------------------------------------------------------------------
// Conv geometry (src/ops/pooling.rs:63-159)
// ---------------------------------------------------------------------------------------
rten_status axis_out(rten_ctx* ctx, int64_t in, int64_t k, int64_t stride, bool same, int64_t ps, int64_t pe,
                     int64_t dil, int64_t* out, int64_t* p0, int64_t* p1) {
    if (dil <= 0) return fail(ctx, RTEN_ERR_INVALID_VALUE, "Dilations must be > 0");
    if (k <= 0) return fail(ctx, RTEN_ERR_INVALID_VALUE, "Kernel size must be > 0");
    if (stride <= 0) return fail(ctx, RTEN_ERR_INVALID_VALUE, "Strides must be > 0");
    if (same) {
        const int64_t o = (in + stride - 1) / stride;
        int64_t total = (o - 1) * stride + (k - 1) * dil + 1 - in;
        if (total < 0) total = 0;
        *out = o;
        *p0 = total / 2;
        *p1 = (total + 1) / 2;
        return RTEN_OK;
    }
    const int64_t padded = in + ps + pe;
    const int64_t dk = k + (k - 1) * (dil - 1);
    if (padded < dk) return fail(ctx, RTEN_ERR_INVALID_VALUE, "Input too small for kernel size");
    *out = (padded - dil * (k - 1) - 1) / stride + 1;
    *p0 = ps;
    *p1 = pe;
    return RTEN_OK;
}

struct ConvArgs {
    int kind;  // 0 f32, 1 int8
    const rten_tensor* x;
    const rten_tensor* w;
    const rten_packed* pw;
    const rten_tensor* bias = nullptr;
    const rten_conv_params* p;
    const rten_tensor* residual = nullptr;
    int act = 0;
    const rten_tensor* x_zp = nullptr;
    const rten_tensor* w_zp = nullptr;
    const rten_tensor* scale = nullptr;
    const rten_tensor* scale_b = nullptr;  // optional second scalar factor (x_scale of a DynamicQuantizeLinear)
    const rten_tensor* out_range = nullptr;  // optional i32[2] device tensor: (min, max) of the output, ordered-int encoded
};

rten_status pack_conv_weight(rten_ctx* ctx, const rten_tensor* w, int esize, void* dst) {
    // OIHW (any strides) -> [O, kh, kw, C]
    const int64_t O = w->shape[0], Cg = w->shape[1], kh = w->shape[2], kw = w->shape[3];
    long long shape[4] = {O, kh, kw, Cg};
    long long ss[4] = {w->strides[0], w->strides[2], w->strides[3], w->strides[1]};
    long long ds[4] = {kh * kw * Cg, kw * Cg, Cg, 1};
    return launch_nd_copy(ctx, esize, w->data, dst, 4, shape, ss, ds);
}

rten_status conv_core(OpScope& sc, ConvArgs& A, rten_tensor* out) {
    rten_ctx* ctx = sc.ctx;
    const rten_conv_params* cp = A.p;
    rten_tensor x, w;
    RTB_TRY(sc.in(A.x, &x));
    RTB_TRY(sc.in(A.w, &w));
    // 1-D convolution via 2-D (conv.rs:142-185)
    const bool one_d = x.ndim == 3;
    int64_t pads_in[4] = {cp->pads[0], cp->pads[1], cp->pads[2], cp->pads[3]};
    int64_t strides[2] = {cp->strides[0], cp->strides[1]}, dil[2] = {cp->dilations[0], cp->dilations[1]};
    if (one_d) {
        if (w.ndim != 3) return fail(ctx, RTEN_ERR_INVALID_VALUE, "input must have 3 dims (OCW)");
        if (cp->n_strides != 1) return fail(ctx, RTEN_ERR_INVALID_VALUE, "expected 1 stride value");
        if (cp->n_dilations != 1) return fail(ctx, RTEN_ERR_INVALID_VALUE, "expected 1 dilation value");
        auto expand = [](rten_tensor& t) {
            t.ndim = 4;
            t.shape[3] = t.shape[2];
            t.strides[3] = t.strides[2];
            t.shape[2] = 1;
            t.strides[2] = 0;
        };
        expand(x);
        expand(w);
        strides[1] = strides[0];
        strides[0] = 1;
        dil[1] = dil[0];
        dil[0] = 1;
        pads_in[1] = cp->pads[0];
        pads_in[3] = cp->pads[1];
        pads_in[0] = pads_in[2] = 0;
    } else {
        if (x.ndim != 4) return fail(ctx, RTEN_ERR_INVALID_VALUE, "input must have 4 dims (NCHW)");
        if (w.ndim != 4) return fail(ctx, RTEN_ERR_INVALID_VALUE, "input must have 4 dims (OCHW)");
        if (cp->n_strides != 2) return fail(ctx, RTEN_ERR_INVALID_VALUE, "expected 2 stride values");
        if (cp->n_dilations != 2) return fail(ctx, RTEN_ERR_INVALID_VALUE, "expected 2 dilation values");
    }
    const int64_t B = x.shape[0], C = x.shape[1], H = x.shape[2], W = x.shape[3];
    const int64_t O = w.shape[0], Cg = w.shape[1], kh = w.shape[2], kw = w.shape[3];
    rten_tensor bias_v;
    if (A.bias) {
        RTB_TRY(sc.in(A.bias, &bias_v));
        if (bias_v.ndim != 1 || bias_v.shape[0] != O)
            return fail(ctx, RTEN_ERR_INCOMPATIBLE_SHAPES, "bias.size(0) != out_channels");
    }
    int64_t OH, OW, pt, pb, pl, pr;
    RTB_TRY(axis_out(ctx, H, kh, strides[0], cp->auto_pad_same != 0, pads_in[0], pads_in[2], dil[0], &OH, &pt, &pb));
    RTB_TRY(axis_out(ctx, W, kw, strides[1], cp->auto_pad_same != 0, pads_in[1], pads_in[3], dil[1], &OW, &pl, &pr));
    const int groups = cp->groups;
    if (groups == 0) return fail(ctx, RTEN_ERR_INVALID_VALUE, "Group count must be > 0");
    if (groups < 0 || C % groups != 0) return fail(ctx, RTEN_ERR_INVALID_VALUE, "Input channel count not divisible by groups");
    if (C / groups != Cg)
        return fail(ctx, RTEN_ERR_INCOMPATIBLE_SHAPES, "Input channels (per group) does not match kernel input channels");
    if (O % groups != 0) return fail(ctx, RTEN_ERR_INVALID_VALUE, "Output channel count not divisible by groups");
    const int64_t Og = O / groups;

    // ---- output (layout follows the input: channels-last in -> channels-last out)
    const int out_dtype = (A.kind == 1 && !A.scale) ? RTEN_I32 : RTEN_F32;
    int64_t oshape[4] = {B, O, OH, OW};
    int64_t pref[4];
    const bool x_cl = (x.strides[1] == 1 && C > 1);
    if (x_cl) {
        pref[1] = 1;
        pref[3] = O;
        pref[2] = OW * O;
        pref[0] = OH * OW * O;
    } else {
        pref[3] = 1;
        pref[2] = OW;
        pref[1] = OH * OW;
        pref[0] = O * OH * OW;
    }
    rten_tensor ov;
    if (one_d) {
        int64_t os3[3] = {B, O, OW};
        int64_t pf3[3] = {pref[0], pref[1], pref[3]};
        RTB_TRY(sc.out(out, out_dtype, 3, os3, &ov, out->data ? nullptr : pf3));
        ov.ndim = 4;
        ov.shape[3] = ov.shape[2];
        ov.strides[3] = ov.strides[2];
        ov.shape[2] = 1;
        ov.strides[2] = 0;
    } else {
        RTB_TRY(sc.out(out, out_dtype, 4, oshape, &ov, out->data ? nullptr : pref));
    }
    if (B * O * OH * OW == 0) return RTEN_OK;

    const int esize = A.kind == 0 ? 4 : 1;
    const int kelems = 128 / esize;

    // ---- weights: prepacked handle or pack per call (the reference prepacks per call too)
    const void* wp = nullptr;
    const int32_t* w_colsum = nullptr;
    if (A.pw) {
        if (A.pw->kind != 1 || A.pw->O != O || A.pw->Cg != Cg || A.pw->kh != kh || A.pw->kw != kw)
            return fail(ctx, RTEN_ERR_INVALID_VALUE, "prepacked conv weight does not match the kernel shape");
        wp = A.pw->data;
        w_colsum = A.pw->colsum;
    } else {
        void* buf = nullptr;
        RTB_TRY(temp_alloc(ctx, (size_t)(O * kh * kw * Cg) * esize, &buf));
        RTB_TRY(pack_conv_weight(ctx, &w, esize, buf));
        wp = buf;
    }

    // ---- integer zero points (x_zp scalar, w_zp per output channel)
    const int32_t* za = nullptr;   // x zero point (GEMM A operand = activations), as i32 ...
    const uint8_t* za8 = nullptr;  // ... or the 8-bit scalar as it is
    const int32_t* zb = nullptr;   // w zero points per column
    int zb_len = 0;
    int pad_value = 0;
    bool x_signed = x.dtype == RTEN_I8, w_signed = w.dtype == RTEN_I8;
    if (A.kind == 1) {
        // padded taps: literal 0 in the reference's shifted-i8 domain (rten-gemm/src/im2col.rs:340-358)
        // == 128 for u8 images, 0 for i8 images
        pad_value = x_signed ? 0 : 128;
        if (A.x_zp) {
            rten_tensor z;
            RTB_TRY(sc.in(A.x_zp, &z));
            if (numel(&z) != 1) return fail(ctx, RTEN_ERR_INVALID_VALUE, "input zero point must be a scalar");
            if (z.dtype != x.dtype) return fail(ctx, RTEN_ERR_CAST_FAILED, "zero point type does not match its tensor");
            za8 = (const uint8_t*)z.data;  // scalar: read in place by the epilogue
            if (!w_colsum) {
                int32_t* cs = nullptr;
                RTB_TRY(temp_alloc(ctx, (size_t)O * 4, (void**)&cs));
                RTB_TRY(launch_rowsum8(ctx, wp, w_signed, O, (int)(kh * kw * Cg), kh * kw * Cg, cs));
                w_colsum = cs;
            }
        }
        if (A.w_zp) {
            RTB_TRY(check_zero_point(ctx, A.w_zp, O, w.dtype));
            rten_tensor z;
            RTB_TRY(sc.in(A.w_zp, &z));
            zb_len = z.ndim == 0 ? 1 : (int)z.shape[0];
            int32_t* p = nullptr;
            RTB_TRY(temp_alloc(ctx, (size_t)zb_len * 4, (void**)&p));
            RTB_TRY(launch_zp_to_i32(ctx, z.data, w_signed, zb_len, z.ndim == 0 ? 0 : z.strides[0], p));
            zb = p;
        }
    }
    const float *scale_p = nullptr, *scale2_p = nullptr;
    if (A.scale) {
        rten_tensor s;
        RTB_TRY(sc.in(A.scale, &s));
        if (numel(&s) != 1 || s.ndim > 1) return fail(ctx, RTEN_ERR_INVALID_VALUE, "scale should be a scalar");
        scale_p = (const float*)s.data;
    }
    if (A.scale_b) {
        rten_tensor s;
        RTB_TRY(sc.in(A.scale_b, &s));
        if (s.dtype != RTEN_F32 || numel(&s) != 1) return fail(ctx, RTEN_ERR_INVALID_VALUE, "scale should be a scalar");
        scale2_p = (const float*)s.data;
    }
    int* range_p = nullptr;
    if (A.out_range) {
        if (A.out_range->dtype != RTEN_I32 || numel(A.out_range) != 2 || A.out_range->device < 0 || !is_contiguous(A.out_range))
            return fail(ctx, RTEN_ERR_INVALID_VALUE, "the output range must be a device-resident i32[2]");
        range_p = (int*)A.out_range->data;
    }
    rten_tensor res_v;
    if (A.residual) {
        RTB_TRY(sc.in(A.residual, &res_v));
        if (res_v.ndim != (one_d ? 3 : 4)) return fail(ctx, RTEN_ERR_INCOMPATIBLE_SHAPES, "residual shape does not match output");
        if (one_d) {
            res_v.ndim = 4;
            res_v.shape[3] = res_v.shape[2];
            res_v.strides[3] = res_v.strides[2];
            res_v.shape[2] = 1;
            res_v.strides[2] = 0;
        }
        for (int i = 0; i < 4; i++)
            if (res_v.shape[i] != oshape[i]) return fail(ctx, RTEN_ERR_INCOMPATIBLE_SHAPES, "residual shape does not match output");
    }

    // ---- choose the addressing path
    //  implicit: x channels-last (c stride 1), C_g*esize % 16 == 0, TMA-addressable strides
    //  explicit: materialise the im2col matrix (odd channel counts such as the 3-channel stem)
    const bool need_pad_copy = (A.kind == 1 && !x_signed && (pt | pb | pl | pr) != 0);
    rten_tensor xs = x;  // source tensor for addressing (may be replaced by an NHWC / padded copy)
    int64_t Hs = H, Ws = W, pt_s = pt, pl_s = pl;
    bool implicit_ok = (Cg * esize) % 16 == 0 && Cg * esize >= 32;
    if (implicit_ok) {
        const bool direct = x.strides[1] == 1 && !need_pad_copy && (reinterpret_cast<uintptr_t>(x.data) % 16 == 0) &&
                            (x.strides[3] * esize) % 16 == 0 && (x.strides[2] * esize) % 16 == 0 &&
                            (x.strides[0] * esize) % 16 == 0;
        if (!direct) {
            // copy to NHWC (with the reference's pad value baked in for u8 images)
            const int64_t Hp = need_pad_copy ? H + pt + pb : H, Wp = need_pad_copy ? W + pl + pr : W;
            void* buf = nullptr;
            RTB_TRY(temp_alloc(ctx, (size_t)(B * Hp * Wp * C) * esize, &buf));
            if (need_pad_copy) RTB_TRY(launch_fill8(ctx, buf, B * Hp * Wp * C, (uint8_t)pad_value));
            long long shape[4] = {B, H, W, C};
            long long ss[4] = {x.strides[0], x.strides[2], x.strides[3], x.strides[1]};
            long long ds[4] = {Hp * Wp * C, Wp * C, C, 1};
            uint8_t* dst = (uint8_t*)buf + (need_pad_copy ? ((pt * Wp + pl) * C) * esize : 0);
            RTB_TRY(launch_nd_copy(ctx, esize, x.data, dst, 4, shape, ss, ds));
            xs.data = buf;
            xs.strides[0] = Hp * Wp * C;
            xs.strides[1] = 1;
            xs.strides[2] = Wp * C;
            xs.strides[3] = C;
            if (need_pad_copy) {
                Hs = Hp;
                Ws = Wp;
                pt_s = 0;
                pl_s = 0;
            }
        }
    }

    // ---- small-channel path (C <= 4, e.g. the RGB stem): NHWC4 zero-padded copy, one 128-byte K block per filter
    //      row holding kw pixels x 4 channels; vertical padding / stride stay in the TMA tile addressing.
    const bool smallc_ok = !implicit_ok && A.kind == 0 && ctx->f32_mode != RTEN_F32_TF32X3 && groups == 1 && Cg <= 4 && kw * 4 <= 32 && dil[1] == 1 &&
                           (int64_t)B * OH * OW > 0 && !getenv("RTEN_B200_NO_SMALLC");
    if (smallc_ok) {
        const int64_t Wp = (OW - 1) * strides[1] + 8;  // every window of 8 pixels stays inside the padded row
        float* xp = nullptr;
        RTB_TRY(temp_alloc(ctx, (size_t)(B * H * Wp * 4) * 4, (void**)&xp));
        RTB_TRY(launch_smallc_pad(ctx, (const float*)x.data, xp, (int)B, (int)C, (int)H, (int)W, (int)Wp, (int)pl, x.strides[0],
                                  x.strides[1], x.strides[2], x.strides[3]));
        float* wsm = nullptr;
        RTB_TRY(temp_alloc(ctx, (size_t)(O * kh * 32) * 4, (void**)&wsm));
        RTB_TRY(launch_smallc_pack_w(ctx, (const float*)w.data, wsm, (int)O, (int)C, (int)kh, (int)kw, w.strides[0], w.strides[1],
                                     w.strides[2], w.strides[3]));
        GemmLaunch L;
        L.kind = 0;
        L.conv = 1;
        L.N = (int)O;
        L.K = (int)(kh * 32);
        L.M = (int)(B * OH * OW);
        L.g.B = (int)B;
        L.g.H = (int)H;
        L.g.W = (int)OW;  // dim 1 of the A map indexes output columns directly
        L.g.C = 32;
        L.g.OH = (int)OH;
        L.g.OW = (int)OW;
        L.g.kh = (int)kh;
        L.g.kw = 1;
        L.g.sy = (int)strides[0];
        L.g.sx = 1;
        L.g.dy = (int)dil[0];
        L.g.dx = 1;
        L.g.pt = (int)pt;
        L.g.pl = 0;
        L.a.base = xp;
        L.a.dims[0] = 32;
        L.a.dims[1] = OW;
        L.a.dims[2] = H;
        L.a.dims[3] = B;
        L.a.strides[0] = 1;
        L.a.strides[1] = strides[1] * 4;
        L.a.strides[2] = Wp * 4;
        L.a.strides[3] = H * Wp * 4;
        L.b.base = wsm;
        L.b.dims[0] = 32;
        L.b.dims[1] = O;
        L.b.dims[2] = kh;
        L.b.dims[3] = 1;
        L.b.strides[0] = 1;
        L.b.strides[1] = kh * 32;
        L.b.strides[2] = 32;
        L.b.strides[3] = 0;
        EpilogueDesc& e = L.epi;
        e.d = ov.data;
        e.s_z0 = ov.strides[0];
        e.s_row = ov.strides[2];
        e.s_z1 = ov.strides[3];
        e.s_col = ov.strides[1];
        e.act = A.act;
        if (A.bias) {
            rten_tensor bc;
            RTB_TRY(sc.contiguous(&bias_v, &bc));
            e.bias = (const float*)bc.data;
            e.bias_kind = 1;
        }
        if (A.residual) {
            e.r = (const float*)res_v.data;
            e.r_scale = 1.0f;
            e.r_z0 = res_v.strides[0];
            e.r_row = res_v.strides[2];
            e.r_z1 = res_v.strides[3];
            e.r_col = res_v.strides[1];
        }
        rten_status st = launch_umma_gemm(ctx, L);
        if (st == RTEN_OK) return RTEN_OK;
        if (st != RTEN_ERR_UNSUPPORTED_VALUE) return st;
        // otherwise fall through to the generic explicit path
    }

    // ---- 8-bit small-channel path (the quantised RGB stem): padded [B,Hp,Wp,16] copy, one 128-byte K block per filter row
    const bool smallc8_ok = !implicit_ok && A.kind == 1 && groups == 1 && Cg <= 16 && kw <= 8 && dil[1] == 1 && !zb &&
                            (int64_t)B * OH * OW > 0 && !getenv("RTEN_B200_NO_SMALLC");
    if (smallc8_ok) {
        const int64_t Hp = H + pt + pb;
        const int64_t Wp = std::max<int64_t>(W + pl + pr, (OW - 1) * strides[1] + 8);
        void *xp = nullptr, *wsm = nullptr;
        RTB_TRY(temp_alloc(ctx, (size_t)(B * Hp * Wp * 16), &xp));
        RTB_TRY(launch_smallc8_pad(ctx, x.data, xp, (int)B, (int)C, (int)H, (int)W, (int)Hp, (int)Wp, (int)pt, (int)pl,
                                   x.strides[0], x.strides[1], x.strides[2], x.strides[3], pad_value));
        RTB_TRY(temp_alloc(ctx, (size_t)(O * kh * 128), &wsm));
        RTB_TRY(launch_smallc8_pack_w(ctx, w.data, wsm, (int)O, (int)C, (int)kh, (int)kw, w.strides[0], w.strides[1],
                                      w.strides[2], w.strides[3]));
        GemmLaunch L;
        L.kind = 1;
        L.a_signed = x_signed;
        L.b_signed = w_signed;
        L.conv = 1;
        L.N = (int)O;
        L.K = (int)(kh * 128);
        L.M = (int)(B * OH * OW);
        L.g.B = (int)B;
        L.g.H = (int)Hp;
        L.g.W = (int)OW;  // dim 1 of the A map indexes output columns directly
        L.g.C = 128;
        L.g.OH = (int)OH;
        L.g.OW = (int)OW;
        L.g.kh = (int)kh;
        L.g.kw = 1;
        L.g.sy = (int)strides[0];
        L.g.sx = 1;
        L.g.dy = (int)dil[0];
        L.g.dx = 1;
        L.g.pt = 0;
        L.g.pl = 0;
        L.a.base = xp;
        L.a.dims[0] = 128;
        L.a.dims[1] = OW;
        L.a.dims[2] = Hp;
        L.a.dims[3] = B;
        L.a.strides[0] = 1;
        L.a.strides[1] = strides[1] * 16;
        L.a.strides[2] = Wp * 16;
        L.a.strides[3] = Hp * Wp * 16;
        L.b.base = wsm;
        L.b.dims[0] = 128;
        L.b.dims[1] = O;
        L.b.dims[2] = kh;
        L.b.dims[3] = 1;
        L.b.strides[0] = 1;
        L.b.strides[1] = kh * 128;
        L.b.strides[2] = 128;
        L.b.strides[3] = 0;
        EpilogueDesc& e = L.epi;
        e.d = ov.data;
        e.d_is_i32 = out_dtype == RTEN_I32;
        e.s_z0 = ov.strides[0];
        e.s_row = ov.strides[2];
        e.s_z1 = ov.strides[3];
        e.s_col = ov.strides[1];
        e.act = A.act;
        if (A.bias) {
            rten_tensor bc;
            RTB_TRY(sc.contiguous(&bias_v, &bc));
            e.bias = (const float*)bc.data;
            e.bias_kind = 1;
        }
        if (A.residual) {
            e.r = (const float*)res_v.data;
            e.r_scale = 1.0f;
            e.r_z0 = res_v.strides[0];
            e.r_row = res_v.strides[2];
            e.r_z1 = res_v.strides[3];
            e.r_col = res_v.strides[1];
        }
        e.za = za;
        e.za_len = za ? 1 : 0;
        e.za8 = za8;
        e.za8_signed = x_signed;
        e.colsum = w_colsum;
        e.scale = scale_p;
        e.scale_len = scale_p ? 1 : 0;
        e.scale2 = scale2_p;
        e.range = range_p;
        rten_status st = launch_umma_gemm(ctx, L);
        if (st == RTEN_OK) return RTEN_OK;
        if (st != RTEN_ERR_UNSUPPORTED_VALUE) return st;
        // otherwise fall through to the generic explicit path
    }

    for (int g = 0; g < groups; g++) {
        GemmLaunch L;
        L.kind = A.kind;
        L.a_signed = x_signed;
        L.b_signed = w_signed;
        L.N = (int)Og;
        L.K = (int)(kh * kw * Cg);
        EpilogueDesc& e = L.epi;
        e.d = (uint8_t*)ov.data + (size_t)(g * Og * ov.strides[1]) * 4;
        e.d_is_i32 = out_dtype == RTEN_I32;
        e.s_z0 = ov.strides[0];
        e.s_row = ov.strides[2];
        e.s_z1 = ov.strides[3];
        e.s_col = ov.strides[1];
        e.act = A.act;
        if (A.bias) {
            rten_tensor bc;
            RTB_TRY(sc.contiguous(&bias_v, &bc));
            e.bias = (const float*)bc.data + g * Og;
            e.bias_kind = 1;
        }
        if (A.residual) {
            e.r = (const float*)res_v.data + g * Og * res_v.strides[1];
            e.r_scale = 1.0f;
            e.r_z0 = res_v.strides[0];
            e.r_row = res_v.strides[2];
            e.r_z1 = res_v.strides[3];
            e.r_col = res_v.strides[1];
        }
        if (A.kind == 1) {
            e.za = za;
            e.za_len = za ? 1 : 0;
            e.za8 = za8;
            e.za8_signed = x_signed;
            e.scale2 = scale2_p;
            e.range = range_p;
            e.colsum = w_colsum ? w_colsum + g * Og : nullptr;
            e.zb = zb ? (zb_len == 1 ? zb : zb + g * Og) : nullptr;
            e.zb_len = zb ? (zb_len == 1 ? 1 : (int)Og) : 0;
            e.scale = scale_p;
            e.scale_len = scale_p ? 1 : 0;
        }
        const uint8_t* wg = (const uint8_t*)wp + (size_t)(g * Og * kh * kw * Cg) * esize;

        if (implicit_ok) {
            L.conv = 1;
            L.g.B = (int)B;
            L.g.H = (int)Hs;
            L.g.W = (int)Ws;
            L.g.C = (int)Cg;
            L.g.OH = (int)OH;
            L.g.OW = (int)OW;
            L.g.kh = (int)kh;
            L.g.kw = (int)kw;
            L.g.sy = (int)strides[0];
            L.g.sx = (int)strides[1];
            L.g.dy = (int)dil[0];
            L.g.dx = (int)dil[1];
            L.g.pt = (int)pt_s;
            L.g.pl = (int)pl_s;
            L.M = (int)(B * OH * OW);
            L.a.base = (const uint8_t*)xs.data + (size_t)(g * Cg) * esize;
            L.a.dims[0] = Cg;
            L.a.dims[1] = Ws;
            L.a.dims[2] = Hs;
            L.a.dims[3] = B;
            L.a.strides[0] = 1;
            L.a.strides[1] = xs.strides[3];
            L.a.strides[2] = xs.strides[2];
            L.a.strides[3] = xs.strides[0];
            L.b.base = wg;
            L.b.dims[0] = Cg;
            L.b.dims[1] = Og;
            L.b.dims[2] = kh * kw;
            L.b.dims[3] = 1;
            L.b.strides[0] = 1;
            L.b.strides[1] = kh * kw * Cg;
            L.b.strides[2] = Cg;
            L.b.strides[3] = 0;
            if (zb) {
                // per-pixel window sums of the image = N=1 GEMM against an all-ones kernel row
                int32_t* rs = nullptr;
                RTB_TRY(temp_alloc(ctx, (size_t)(B * OH * OW) * 4, (void**)&rs));
                void* ones = nullptr;
                RTB_TRY(temp_alloc(ctx, (size_t)(kh * kw * Cg), &ones));
                RTB_TRY(launch_fill8(ctx, ones, kh * kw * Cg, 1));
                GemmLaunch R = L;
                R.N = 1;
                R.b_signed = 1;
                R.b.base = ones;
                R.b.dims[1] = 1;
                R.epi = EpilogueDesc();
                R.epi.d = rs;
                R.epi.d_is_i32 = 1;
                R.epi.s_z0 = OH * OW;
                R.epi.s_row = OW;
                R.epi.s_z1 = 1;
                R.epi.s_col = B * OH * OW;
                rten_status st = launch_umma_gemm(ctx, R);
                if (st != RTEN_OK) return fail(ctx, st, "conv window-sum GEMM could not be launched");
                e.rowsum = rs;
            }
            rten_status st = launch_umma_gemm(ctx, L);
            if (st == RTEN_OK) continue;
            if (st != RTEN_ERR_UNSUPPORTED_VALUE) return st;
            // fall through to the explicit path
        }
        // explicit im2col: A = [B*OH*OW, kpad]
        const int64_t Kd = kh * kw * Cg;
        const int64_t kpad = round_up(Kd, 16 / esize);
        void* col = nullptr;
        RTB_TRY(temp_alloc(ctx, (size_t)(B * OH * OW * kpad) * esize, &col));
        Im2ColParams ip;
        ip.B = (int)B;
        ip.C = (int)Cg;
        ip.H = (int)H;
        ip.W = (int)W;
        ip.OH = (int)OH;
        ip.OW = (int)OW;
        ip.kh = (int)kh;
        ip.kw = (int)kw;
        ip.sy = (int)strides[0];
        ip.sx = (int)strides[1];
        ip.dy = (int)dil[0];
        ip.dx = (int)dil[1];
        ip.pt = (int)pt;
        ip.pl = (int)pl;
        ip.c0 = (int)(g * Cg);
        ip.kpad = (int)kpad;
        ip.xs_b = x.strides[0];
        ip.xs_c = x.strides[1];
        ip.xs_h = x.strides[2];
        ip.xs_w = x.strides[3];
        RTB_TRY(launch_im2col(ctx, esize, x.data, col, ip, pad_value));
        L.conv = 0;
        L.M = (int)(B * OH * OW);
        L.z0 = L.z1 = 1;
        L.a = OperandDesc();
        L.a.base = col;
        L.a.dims[0] = Kd;
        L.a.dims[1] = B * OH * OW;
        L.a.strides[1] = kpad;
        L.b = OperandDesc();
        L.b.base = wg;
        L.b.dims[0] = Kd;
        L.b.dims[1] = Og;
        L.b.strides[1] = Kd;
        if (!tma_compatible(L.b, esize, 4)) {
            void* wb = nullptr;
            RTB_TRY(temp_alloc(ctx, (size_t)(Og * kpad) * esize, &wb));
            long long shape[2] = {Og, Kd}, ss[2] = {Kd, 1}, ds[2] = {kpad, 1};
            RTB_TRY(launch_nd_copy(ctx, esize, wg, wb, 2, shape, ss, ds));
            L.b.base = wb;
            L.b.strides[1] = kpad;
        }
        // plain-mode epilogue needs a uniform row stride over (b, oy, ox): use the conv decomposition when the
        // output is not pixel-contiguous by writing through a temp
        const bool uniform = (ov.strides[2] == OW * ov.strides[3]) && (ov.strides[0] == OH * ov.strides[2] || B == 1);
        const bool r_uniform = !A.residual || ((res_v.strides[2] == OW * res_v.strides[3]) &&
                                               (res_v.strides[0] == OH * res_v.strides[2] || B == 1));
        if (zb) {
            int32_t* rs = nullptr;
            RTB_TRY(temp_alloc(ctx, (size_t)(B * OH * OW) * 4, (void**)&rs));
            RTB_TRY(launch_rowsum8(ctx, col, x_signed, B * OH * OW, (int)Kd, kpad, rs));
            e.rowsum = rs;
        }
        if (uniform && r_uniform) {
            e.s_row = ov.strides[3];
            e.s_z0 = e.s_z1 = 0;
            if (A.residual) {
                e.r_row = res_v.strides[3];
                e.r_z0 = e.r_z1 = 0;
            }
            rten_status st = launch_umma_gemm(ctx, L);
            if (st != RTEN_OK) return fail(ctx, st, "conv GEMM could not be launched");
        } else {
            // NCHW-style output: GEMM into [pixels, Og] temp (no fusion), then strided copy + residual/act
            void* tmp = nullptr;
            RTB_TRY(temp_alloc(ctx, (size_t)(B * OH * OW * Og) * 4, &tmp));
            EpilogueDesc e2 = e;
            e2.d = tmp;
            e2.s_row = Og;
            e2.s_col = 1;
            e2.s_z0 = e2.s_z1 = 0;
            e2.r = nullptr;
            e2.act = A.residual ? 0 : A.act;
            GemmLaunch L2 = L;
            L2.epi = e2;
            rten_status st = launch_umma_gemm(ctx, L2);
            if (st != RTEN_OK) return fail(ctx, st, "conv GEMM could not be launched");
            long long shape[4] = {B, OH, OW, Og};
            long long ss[4] = {OH * OW * Og, OW * Og, Og, 1};
            long long ds[4] = {ov.strides[0], ov.strides[2], ov.strides[3], ov.strides[1]};
            if (A.residual) {
                long long rs4[4] = {res_v.strides[0], res_v.strides[2], res_v.strides[3], res_v.strides[1]};
                RTB_TRY(launch_nd_add(ctx, (const float*)tmp, e.r, (float*)e.d, 4, shape, ss, rs4, ds, A.act == 1));
                if (A.act > 1) return fail(ctx, RTEN_ERR_UNSUPPORTED_VALUE, "gelu after residual needs a pixel-contiguous output");
            } else {
                RTB_TRY(launch_nd_copy(ctx, 4, tmp, e.d, 4, shape, ss, ds));
            }
        }
    }
    return RTEN_OK;
}

}  // namespace

// =========================================================================================
extern "C" {

rten_status rten_b200_prepack_b(rten_ctx* ctx, const rten_tensor* b, rten_packed** out) {
    RTB_TRY(check_ctx(ctx));
    if (!b || !out) return RTEN_ERR_INVALID_VALUE;
    *out = nullptr;
    if (b->ndim != 2) return fail(ctx, RTEN_ERR_INVALID_VALUE, "prepack expects a matrix");  // matmul_prepack_b: try_into Matrix
    if (b->dtype == RTEN_I32) return fail(ctx, RTEN_ERR_UNSUPPORTED_TYPE, "unsupported type");
    OpScope sc(ctx);
    rten_tensor bv;
    rten_status st = sc.in(b, &bv);
    rten_packed* p = nullptr;
    if (st == RTEN_OK) {
        const int es = dtype_size(b->dtype);
        p = new rten_packed();
        p->kind = 0;
        p->dtype = b->dtype;
        p->K = bv.shape[0];
        p->N = bv.shape[1];
        p->ld = round_up(std::max<int64_t>(p->K, 1), 16 / es);
        st = pool_alloc(ctx, (size_t)std::max<int64_t>(p->N * p->ld, 1) * es, &p->data);
        if (st == RTEN_OK) {
            RTB_CUDA(ctx, cudaMemsetAsync(p->data, 0, (size_t)std::max<int64_t>(p->N * p->ld, 1) * es, rtb::launch_stream(ctx)));
            long long shape[2] = {p->N, p->K}, ss[2] = {bv.strides[1], bv.strides[0]}, ds[2] = {p->ld, 1};
            st = launch_nd_copy(ctx, es, bv.data, p->data, 2, shape, ss, ds);
        }
        if (st == RTEN_OK && es == 1 && p->N > 0) {
            st = pool_alloc(ctx, (size_t)p->N * 4, (void**)&p->colsum);
            if (st == RTEN_OK) st = launch_rowsum8(ctx, p->data, p->dtype == RTEN_I8, p->N, (int)p->K, p->ld, p->colsum);
        }
    }
    st = sc.finish(st);
    if (st != RTEN_OK) {
        if (p) rten_b200_packed_free(ctx, p);
        return st;
    }
    *out = p;
    return RTEN_OK;
}

rten_status rten_b200_prepack_conv_weight(rten_ctx* ctx, const rten_tensor* w, int groups, rten_packed** out) {
    RTB_TRY(check_ctx(ctx));
    if (!w || !out) return RTEN_ERR_INVALID_VALUE;
    *out = nullptr;
    if (w->ndim != 4) return fail(ctx, RTEN_ERR_INVALID_VALUE, "input must have 4 dims (OCHW)");
    if (w->dtype == RTEN_I32) return fail(ctx, RTEN_ERR_UNSUPPORTED_TYPE, "unsupported type");
    OpScope sc(ctx);
    rten_tensor wv;
    rten_status st = sc.in(w, &wv);
    rten_packed* p = nullptr;
    if (st == RTEN_OK) {
        const int es = dtype_size(w->dtype);
        p = new rten_packed();
        p->kind = 1;
        p->dtype = w->dtype;
        p->O = wv.shape[0];
        p->Cg = wv.shape[1];
        p->kh = wv.shape[2];
        p->kw = wv.shape[3];
        p->groups = groups;
        const int64_t n = std::max<int64_t>(p->O * p->Cg * p->kh * p->kw, 1);
        st = pool_alloc(ctx, (size_t)n * es, &p->data);
        if (st == RTEN_OK) st = pack_conv_weight(ctx, &wv, es, p->data);
        if (st == RTEN_OK && es == 1 && p->O > 0) {
            st = pool_alloc(ctx, (size_t)p->O * 4, (void**)&p->colsum);
            const int64_t Kd = p->Cg * p->kh * p->kw;
            if (st == RTEN_OK) st = launch_rowsum8(ctx, p->data, p->dtype == RTEN_I8, p->O, (int)Kd, Kd, p->colsum);
        }
    }
    st = sc.finish(st);
    if (st != RTEN_OK) {
        if (p) rten_b200_packed_free(ctx, p);
        return st;
    }
    *out = p;
    return RTEN_OK;
}

void rten_b200_packed_free(rten_ctx* ctx, rten_packed* p) {
    if (!p) return;
    if (ctx) {
        pool_free(ctx, p->data);
        pool_free(ctx, p->colsum);
    }
    delete p;
}

// ---- Gemm ---------------------------------------------------------------------------------
rten_status rten_b200_gemm(rten_ctx* ctx, const rten_tensor* a, const rten_tensor* b, const rten_tensor* c, float alpha,
                           float beta, int trans_a, int trans_b, rten_tensor* out) {
    RTB_TRY(check_ctx(ctx));
    if (!a || !b || !out) return fail(ctx, RTEN_ERR_MISSING_INPUTS, "missing inputs");
    if (a->dtype != RTEN_F32 || b->dtype != RTEN_F32 || (c && c->dtype != RTEN_F32))
        return fail(ctx, RTEN_ERR_UNSUPPORTED_TYPE, "unsupported type");
    if (a->ndim != 2 || b->ndim != 2) return fail(ctx, RTEN_ERR_INVALID_VALUE, "input must have 2 dims");
    OpScope sc(ctx);
    rten_tensor av, bv, cv;
    rten_status st = sc.in(a, &av);
    if (st == RTEN_OK) st = sc.in(b, &bv);
    if (st == RTEN_OK && c) st = sc.in(c, &cv);
    if (st == RTEN_OK) {
        auto transpose = [](rten_tensor& t) {
            std::swap(t.shape[0], t.shape[1]);
            std::swap(t.strides[0], t.strides[1]);
        };
        if (trans_a) transpose(av);
        if (trans_b) transpose(bv);
        if (av.shape[1] != bv.shape[0]) {
            st = fail(ctx, RTEN_ERR_INCOMPATIBLE_SHAPES, "Columns of first matrix does not match rows of second matrix");
        } else {
            MatMulArgs A{};
            A.kind = 0;
            A.a = &av;
            A.b = &bv;
            A.pb = nullptr;
            A.out_dtype = RTEN_F32;
            A.epi.alpha = alpha;
            const int64_t M = av.shape[0], N = bv.shape[1];
            if (c && beta != 0.0f) {
                // broadcast c to [M, N] (matmul.rs:63-67)
                int64_t cs[2] = {0, 0};
                bool ok = cv.ndim <= 2;
                if (ok) {
                    for (int i = 0; i < cv.ndim; i++) {
                        const int od = 2 - cv.ndim + i;
                        const int64_t want = od == 0 ? M : N;
                        if (cv.shape[i] == want)
                            cs[od] = cv.strides[i];
                        else if (cv.shape[i] == 1)
                            cs[od] = 0;
                        else
                            ok = false;
                    }
                }
                if (!ok) {
                    st = fail(ctx, RTEN_ERR_INCOMPATIBLE_SHAPES, "Cannot broadcast c to output shape");
                } else {
                    A.epi.r = (const float*)cv.data;
                    A.epi.r_scale = beta;
                    A.epi.r_row = cs[0];
                    A.epi.r_col = cs[1];
                }
            }
            if (st == RTEN_OK) {
                // matmul_core's residual plumbing is for same-shape tensors; C is already set in epi.
                st = matmul_core(sc, A, out);
            }
        }
    }
    return sc.finish(st);
}

// ---- MatMul / FusedMatMul -----------------------------------------------------------------
rten_status rten_b200_matmul_ex(rten_ctx* ctx, const rten_tensor* a, const rten_tensor* b, const rten_packed* pb,
                                const rten_tensor* bias, float alpha, const rten_tensor* residual, int activation,
                                rten_tensor* out) {
    RTB_TRY(check_ctx(ctx));
    if (!a || !b || !out) return fail(ctx, RTEN_ERR_MISSING_INPUTS, "missing inputs");
    if (a->dtype != RTEN_F32 || b->dtype != RTEN_F32 || (pb && pb->dtype != RTEN_F32))
        return fail(ctx, RTEN_ERR_UNSUPPORTED_TYPE, "unsupported type");
    OpScope sc(ctx);
    rten_tensor av, bv, biasv, biasc;
    rten_status st = sc.in(a, &av);
    if (st == RTEN_OK) {
        if (pb) {
            bv = *b;  // only the shape is consulted
            bv.data = nullptr;
        } else {
            st = sc.in(b, &bv);
        }
    }
    MatMulArgs A{};
    A.kind = 0;
    A.a = &av;
    A.b = &bv;
    A.pb = pb;
    A.out_dtype = RTEN_F32;
    A.epi.alpha = alpha;
    A.epi.act = activation;
    A.residual = residual;
    if (st == RTEN_OK && bias) {
        if (bias->dtype != RTEN_F32 || bias->ndim != 1) {
            st = fail(ctx, RTEN_ERR_CAST_FAILED, "bias must be a float vector");
        } else {
            st = sc.in(bias, &biasv);
            if (st == RTEN_OK) st = sc.contiguous(&biasv, &biasc);
            const int64_t N = bv.ndim >= 2 ? bv.shape[bv.ndim - 1] : 1;
            if (st == RTEN_OK && biasc.shape[0] != N) st = fail(ctx, RTEN_ERR_INVALID_VALUE, "WrongBiasSize");
            A.epi.bias = (const float*)biasc.data;
            A.epi.bias_kind = 1;
        }
    }
    if (st == RTEN_OK) st = matmul_core(sc, A, out);
    return sc.finish(st);
}

rten_status rten_b200_matmul(rten_ctx* ctx, const rten_tensor* a, const rten_tensor* b, const rten_packed* pb,
                             const rten_tensor* bias, float alpha, rten_tensor* out) {
    return rten_b200_matmul_ex(ctx, a, b, pb, bias, alpha, nullptr, 0, out);
}

// ---- MatMulInteger / MatMulIntegerToFloat ------------------------------------------------------
rten_status rten_b200_matmul_integer(rten_ctx* ctx, const rten_tensor* a, const rten_tensor* b, const rten_packed* pb,
                                     const rten_tensor* a_zp, const rten_tensor* b_zp, const rten_tensor* scale,
                                     rten_tensor* out) {
    return rten_b200_matmul_integer_ex(ctx, a, b, pb, a_zp, b_zp, scale, nullptr, nullptr, nullptr, 0, nullptr, out);
}

rten_status rten_b200_matmul_integer_ex(rten_ctx* ctx, const rten_tensor* a, const rten_tensor* b, const rten_packed* pb,
                                        const rten_tensor* a_zp, const rten_tensor* b_zp, const rten_tensor* scale,
                                        const rten_tensor* scale_b, const rten_tensor* bias, const rten_tensor* residual,
                                        int activation, rten_tensor* out_range, rten_tensor* out) {
    RTB_TRY(check_ctx(ctx));
    if (!a || !b || !out) return fail(ctx, RTEN_ERR_MISSING_INPUTS, "missing inputs");
    if ((bias || residual || activation || scale_b) && !scale)
        return fail(ctx, RTEN_ERR_INVALID_VALUE, "bias / residual / activation follow the float conversion: a scale is required");
    if (activation < 0 || activation > 3) return fail(ctx, RTEN_ERR_INVALID_VALUE, "unknown activation");
    auto is8 = [](int dt) { return dt == RTEN_U8 || dt == RTEN_I8; };
    if (!is8(a->dtype) || !is8(b->dtype)) return fail(ctx, RTEN_ERR_UNSUPPORTED_TYPE, "unsupported type");
    const int64_t a_rows = a->ndim > 1 ? a->shape[a->ndim - 2] : 1;
    const int64_t b_cols = b->ndim > 1 ? b->shape[b->ndim - 1] : 1;
    RTB_TRY(check_zero_point(ctx, a_zp, a_rows, a->dtype));
    RTB_TRY(check_zero_point(ctx, b_zp, b_cols, b->dtype));
    OpScope sc(ctx);
    rten_tensor av, bv, sv, svc;
    rten_status st = sc.in(a, &av);
    if (st == RTEN_OK) {
        if (pb && pb->dtype == b->dtype) {
            bv = *b;
            bv.data = nullptr;
        } else {
            pb = nullptr;
            st = sc.in(b, &bv);
        }
    }
    MatMulArgs A{};
    A.kind = 1;
    A.a = &av;
    A.b = &bv;
    A.pb = pb;
    A.a_zp = a_zp;
    A.b_zp = b_zp;
    A.out_dtype = scale ? RTEN_F32 : RTEN_I32;
    if (st == RTEN_OK && scale) {
        // OutputScale::from_view (matmul.rs:712-721)
        if (scale->dtype != RTEN_F32) {
            st = fail(ctx, RTEN_ERR_CAST_FAILED, "scale must be float");
        } else if (scale->ndim > 1) {
            st = fail(ctx, RTEN_ERR_INVALID_VALUE, "scale should have rank 0 or 1");
        } else {
            st = sc.in(scale, &sv);
            if (st == RTEN_OK) st = sc.contiguous(&sv, &svc);
            const int64_t len = svc.ndim == 0 ? 1 : svc.shape[0];
            if (st == RTEN_OK && len != 1 && len != b_cols)
                st = fail(ctx, RTEN_ERR_INCOMPATIBLE_SHAPES, "Scale length does not match tensor columns");
            A.epi.scale = (const float*)svc.data;
            A.epi.scale_len = (int)len;
        }
    }
    rten_tensor biasv, biasc, s2v;
    if (st == RTEN_OK && scale_b) {
        if (scale_b->dtype != RTEN_F32 || numel(scale_b) != 1) {
            st = fail(ctx, RTEN_ERR_INVALID_VALUE, "the second scale factor must be a float scalar");
        } else {
            st = sc.in(scale_b, &s2v);
            A.epi.scale2 = (const float*)s2v.data;
        }
    }
    if (st == RTEN_OK && bias) {
        if (bias->dtype != RTEN_F32 || bias->ndim != 1) {
            st = fail(ctx, RTEN_ERR_CAST_FAILED, "bias must be a float vector");
        } else {
            st = sc.in(bias, &biasv);
            if (st == RTEN_OK) st = sc.contiguous(&biasv, &biasc);
            if (st == RTEN_OK && biasc.shape[0] != b_cols) st = fail(ctx, RTEN_ERR_INVALID_VALUE, "WrongBiasSize");
            A.epi.bias = (const float*)biasc.data;
            A.epi.bias_kind = 1;
        }
    }
    A.epi.act = activation;
    A.residual = residual;
    if (st == RTEN_OK && out_range) {
        if (!scale || out_range->dtype != RTEN_I32 || numel(out_range) != 2 || out_range->device < 0 || !is_contiguous(out_range))
            st = fail(ctx, RTEN_ERR_INVALID_VALUE, "the output range must be a device-resident i32[2] (float outputs only)");
        else
            A.epi.range = (int*)out_range->data;
    }
    if (st == RTEN_OK) st = matmul_core(sc, A, out);
    return sc.finish(st);
}

// ---- Conv family ----------------------------------------------------------------------------
rten_status rten_b200_conv2d_ex(rten_ctx* ctx, const rten_tensor* x, const rten_tensor* w, const rten_packed* pw,
                                const rten_tensor* bias, const rten_conv_params* p, const rten_tensor* residual,
                                int activation, rten_tensor* out) {
    RTB_TRY(check_ctx(ctx));
    if (!x || !w || !p || !out) return fail(ctx, RTEN_ERR_MISSING_INPUTS, "missing inputs");
    if (x->dtype != RTEN_F32 || w->dtype != RTEN_F32 || (bias && bias->dtype != RTEN_F32))
        return fail(ctx, RTEN_ERR_UNSUPPORTED_TYPE, "unsupported type");
    OpScope sc(ctx);
    ConvArgs A{};
    A.kind = 0;
    A.x = x;
    A.w = w;
    A.pw = pw;
    A.bias = bias;
    A.p = p;
    A.residual = residual;
    A.act = activation;
    return sc.finish(conv_core(sc, A, out));
}

rten_status rten_b200_conv2d(rten_ctx* ctx, const rten_tensor* x, const rten_tensor* w, const rten_packed* pw,
                             const rten_tensor* bias, const rten_conv_params* p, rten_tensor* out) {
    return rten_b200_conv2d_ex(ctx, x, w, pw, bias, p, nullptr, 0, out);
}

rten_status rten_b200_conv_integer(rten_ctx* ctx, const rten_tensor* x, const rten_tensor* w, const rten_packed* pw,
                                   const rten_tensor* x_zp, const rten_tensor* w_zp, const rten_tensor* scale,
                                   const rten_conv_params* p, rten_tensor* out) {
    return rten_b200_conv_integer_ex(ctx, x, w, pw, x_zp, w_zp, scale, nullptr, p, nullptr, nullptr, 0, nullptr, out);
}

rten_status rten_b200_conv_integer_ex(rten_ctx* ctx, const rten_tensor* x, const rten_tensor* w, const rten_packed* pw,
                                      const rten_tensor* x_zp, const rten_tensor* w_zp, const rten_tensor* scale,
                                      const rten_tensor* scale_b, const rten_conv_params* p, const rten_tensor* bias,
                                      const rten_tensor* residual, int activation, rten_tensor* out_range,
                                      rten_tensor* out) {
    RTB_TRY(check_ctx(ctx));
    if (!x || !w || !p || !out) return fail(ctx, RTEN_ERR_MISSING_INPUTS, "missing inputs");
    auto is8 = [](int dt) { return dt == RTEN_U8 || dt == RTEN_I8; };
    if (!is8(x->dtype) || !is8(w->dtype)) return fail(ctx, RTEN_ERR_UNSUPPORTED_TYPE, "unsupported type");
    if (scale && scale->dtype != RTEN_F32) return fail(ctx, RTEN_ERR_CAST_FAILED, "scale must be float");
    if ((bias || residual || activation || scale_b) && !scale)
        return fail(ctx, RTEN_ERR_INVALID_VALUE, "bias / residual / activation follow the float conversion: a scale is required");
    if ((bias && bias->dtype != RTEN_F32) || (residual && residual->dtype != RTEN_F32))
        return fail(ctx, RTEN_ERR_UNSUPPORTED_TYPE, "unsupported type");
    if (activation < 0 || activation > 1) return fail(ctx, RTEN_ERR_UNSUPPORTED_VALUE, "only Relu can follow an integer convolution");
    OpScope sc(ctx);
    ConvArgs A{};
    A.kind = 1;
    A.x = x;
    A.w = w;
    A.pw = (pw && pw->dtype == w->dtype) ? pw : nullptr;
    A.p = p;
    A.x_zp = x_zp;
    A.w_zp = w_zp;
    A.scale = scale;
    A.scale_b = scale_b;
    A.bias = bias;
    A.residual = residual;
    A.act = activation;
    A.out_range = out_range;
    if (out_range && !scale) return fail(ctx, RTEN_ERR_INVALID_VALUE, "the output range is defined for float outputs: a scale is required");
    return sc.finish(conv_core(sc, A, out));
}

// ---- Softmax / AddSoftmax -------------------------------------------------------------------
rten_status rten_b200_softmax(rten_ctx* ctx, const rten_tensor* x, const rten_tensor* mask, int axis, int flush_nans,
                              rten_tensor* out) {
    RTB_TRY(check_ctx(ctx));
    if (!x || !out) return fail(ctx, RTEN_ERR_MISSING_INPUTS, "missing inputs");
    if (x->dtype != RTEN_F32 || (mask && mask->dtype != RTEN_F32)) return fail(ctx, RTEN_ERR_UNSUPPORTED_TYPE, "unsupported type");
    const int nd = x->ndim;
    if (nd == 0 || axis < -nd || axis >= nd) return fail(ctx, RTEN_ERR_INVALID_VALUE, "Axis is invalid");
    const int ax = axis < 0 ? axis + nd : axis;
    OpScope sc(ctx);
    rten_tensor xv, mv, ov;
    rten_status st = sc.in(x, &xv);
    if (st == RTEN_OK && mask) st = sc.in(mask, &mv);
    if (st == RTEN_OK) st = sc.out(out, RTEN_F32, nd, xv.shape, &ov, nullptr);
    if (st == RTEN_OK && numel(&xv) > 0) {
        // view with `ax` moved last
        int perm[RTEN_MAX_DIMS], k = 0;
        for (int i = 0; i < nd; i++)
            if (i != ax) perm[k++] = i;
        perm[nd - 1] = ax;
        rten_tensor xp = xv, op = ov;
        for (int i = 0; i < nd; i++) {
            xp.shape[i] = xv.shape[perm[i]];
            xp.strides[i] = xv.strides[perm[i]];
            op.shape[i] = ov.shape[perm[i]];
            op.strides[i] = ov.strides[perm[i]];
        }
        rten_tensor xc;
        st = sc.contiguous(&xp, &xc);
        // run in place on the output when it is lane-contiguous in the permuted view, else via temp
        rten_tensor yc = op;
        const bool out_direct = is_contiguous(&op);
        if (st == RTEN_OK && !out_direct) {
            set_contiguous(&yc);
            void* t = nullptr;
            st = temp_alloc(ctx, (size_t)numel(&xv) * 4, &t);
            yc.data = t;
        }
        if (st == RTEN_OK) {
            const int n = (int)xp.shape[nd - 1];
            const long long rows = numel(&xv) / n;
            const float* mp = nullptr;
            long long lead[4] = {1, 1, 1, 1}, ms[4] = {0, 0, 0, 0}, ms_last = 0;
            int nlead = 0;
            if (mask) {
                // broadcast mask to x's shape (numpy rules), in the permuted dim order
                if (mv.ndim > nd) {
                    st = fail(ctx, RTEN_ERR_INCOMPATIBLE_SHAPES, "Cannot broadcast inputs");
                } else {
                    long long mstr[RTEN_MAX_DIMS];
                    for (int i = 0; i < nd && st == RTEN_OK; i++) {
                        const int mi = i - (nd - mv.ndim);
                        if (mi < 0 || mv.shape[mi] == 1)
                            mstr[i] = 0;
                        else if (mv.shape[mi] == xv.shape[i])
                            mstr[i] = mv.strides[mi];
                        else
                            st = fail(ctx, RTEN_ERR_INCOMPATIBLE_SHAPES, "Cannot broadcast inputs");
                    }
                    if (st == RTEN_OK) {
                        // leading dims in permuted order, collapsed where the mask advances uniformly
                        std::vector<long long> ls, lst;
                        for (int i = 0; i < nd - 1; i++) {
                            const long long s = xp.shape[i], stv = mstr[perm[i]];
                            if (s == 1) continue;
                            if (!ls.empty() && lst.back() == stv * s) {
                                ls.back() *= s;
                                lst.back() = stv;
                            } else {
                                ls.push_back(s);
                                lst.push_back(stv);
                            }
                        }
                        if (ls.size() > 4) {
                            st = fail(ctx, RTEN_ERR_UNSUPPORTED_VALUE, "mask broadcast pattern needs more than 4 strided dims");
                        } else {
                            nlead = (int)ls.size();
                            for (int i = 0; i < nlead; i++) {
                                lead[i] = ls[i];
                                ms[i] = lst[i];
                            }
                            ms_last = mstr[ax];
                            mp = (const float*)mv.data;
                        }
                    }
                }
            }
            if (st == RTEN_OK)
                st = launch_softmax(ctx, (const float*)xc.data, (float*)yc.data, rows, n, flush_nans, mp, nlead, lead, ms,
                                    ms_last);
            if (st == RTEN_OK && !out_direct) {
                long long shape[RTEN_MAX_DIMS], ss[RTEN_MAX_DIMS], ds[RTEN_MAX_DIMS];
                for (int i = 0; i < nd; i++) {
                    shape[i] = yc.shape[i];
                    ss[i] = yc.strides[i];
                    ds[i] = op.strides[i];
                }
                st = launch_nd_copy(ctx, 4, yc.data, op.data, nd, shape, ss, ds);
            }
        }
    }
    return sc.finish(st);
}

// ---- LayerNormalization -------------------------------------------------------------------------
rten_status rten_b200_layer_norm(rten_ctx* ctx, const rten_tensor* x, const rten_tensor* scale, const rten_tensor* bias,
                                 int axis, float epsilon, rten_tensor* out) {
    RTB_TRY(check_ctx(ctx));
    if (!x || !scale || !out) return fail(ctx, RTEN_ERR_MISSING_INPUTS, "missing inputs");
    if (x->dtype != RTEN_F32 || scale->dtype != RTEN_F32 || (bias && bias->dtype != RTEN_F32))
        return fail(ctx, RTEN_ERR_UNSUPPORTED_TYPE, "unsupported type");
    const int nd = x->ndim;
    if (axis < -nd || axis >= std::max(nd, 1)) return fail(ctx, RTEN_ERR_INVALID_VALUE, "Axis is invalid");
    const int ax = axis < 0 ? axis + nd : axis;
    const float eps = epsilon < 0.0f ? 1e-5f : epsilon;
    OpScope sc(ctx);
    rten_tensor xv, xc, sv, bv, ov;
    rten_status st = sc.in(x, &xv);
    if (st == RTEN_OK) st = sc.in(scale, &sv);
    if (st == RTEN_OK && bias) st = sc.in(bias, &bv);
    const int nn = nd - ax;  // normalized dims
    // broadcast a parameter to the normalized shape, materialised contiguous; scalars stay scalar
    auto param = [&](const rten_tensor& pv, const char* err, const float** ptr, float* scalar, bool* is_scalar) -> rten_status {
        if (numel(&pv) == 1) {
            // scale.item(): read on device later -> use a 1-element broadcast with stride 0
            *is_scalar = true;
            *ptr = (const float*)pv.data;
            (void)scalar;
            return RTEN_OK;
        }
        *is_scalar = false;
        if (pv.ndim > nn) return fail(ctx, RTEN_ERR_INVALID_VALUE, err);
        rten_tensor b = pv;
        b.ndim = nn;
        for (int i = 0; i < nn; i++) {
            const int pi = i - (nn - pv.ndim);
            const int64_t want = xv.shape[ax + i];
            b.shape[i] = want;
            if (pi < 0 || pv.shape[pi] == 1)
                b.strides[i] = 0;
            else if (pv.shape[pi] == want)
                b.strides[i] = pv.strides[pi];
            else
                return fail(ctx, RTEN_ERR_INVALID_VALUE, err);
        }
        rten_tensor c;
        RTB_TRY(sc.contiguous(&b, &c));
        *ptr = (const float*)c.data;
        return RTEN_OK;
    };
    const float *gp = nullptr, *bp = nullptr;
    float gs = 1.0f, bs = 0.0f;
    bool g_scalar = false, b_scalar = false;
    if (st == RTEN_OK) st = param(sv, "`scale` is not broadcastable to normalized axes of input", &gp, &gs, &g_scalar);
    if (st == RTEN_OK && bias) st = param(bv, "`bias` is not broadcastable to normalized axes of input", &bp, &bs, &b_scalar);
    if (st == RTEN_OK) st = sc.contiguous(&xv, &xc);
    if (st == RTEN_OK) st = sc.out(out, RTEN_F32, nd, xv.shape, &ov, nullptr);
    if (st == RTEN_OK && numel(&xv) > 0) {
        long long n = 1;
        for (int i = ax; i < nd; i++) n *= xv.shape[i];
        const long long rows = numel(&xv) / n;
        // scalar gamma / beta stay on the device and are read by the kernel (the reference's scalar-scale arm computes
        // rstd = scale / sqrt(var + eps), src/ops/norm.rs:456-529): no host read, no synchronisation, capturable
        const float *gsp = nullptr, *bsp = nullptr;
        if (g_scalar) {
            gsp = gp;
            gp = nullptr;
        }
        if (bias && b_scalar) {
            bsp = bp;
            bp = nullptr;
        }
        rten_tensor yc = ov;
        const bool direct = is_contiguous(&ov);
        if (!direct) {
            set_contiguous(&yc);
            void* t = nullptr;
            st = temp_alloc(ctx, (size_t)numel(&xv) * 4, &t);
            yc.data = t;
        }
        if (st == RTEN_OK)
            st = launch_layer_norm(ctx, (const float*)xc.data, (float*)yc.data, rows, (int)n, gp, gs, bp, bs, eps, gsp, bsp);
        if (st == RTEN_OK && !direct) {
            long long shape[RTEN_MAX_DIMS], ss[RTEN_MAX_DIMS], ds[RTEN_MAX_DIMS];
            for (int i = 0; i < nd; i++) {
                shape[i] = yc.shape[i];
                ss[i] = yc.strides[i];
                ds[i] = ov.strides[i];
            }
            st = launch_nd_copy(ctx, 4, yc.data, ov.data, nd, shape, ss, ds);
        }
    }
    return sc.finish(st);
}

// ---- unary elementwise ----------------------------------------------------------------------------
static rten_status unary_op(rten_ctx* ctx, int op, const rten_tensor* x, rten_tensor* out) {
    RTB_TRY(check_ctx(ctx));
    if (!x || !out) return fail(ctx, RTEN_ERR_MISSING_INPUTS, "missing inputs");
    if (x->dtype != RTEN_F32) return fail(ctx, RTEN_ERR_UNSUPPORTED_TYPE, "unsupported type");
    OpScope sc(ctx);
    rten_tensor xv, ov;
    rten_status st = sc.in(x, &xv);
    // dense (any dim order) tensors are processed in memory order: output takes the input's strides
    bool dense = false;
    if (st == RTEN_OK) dense = span_elems(&xv) == numel(&xv);
    if (st == RTEN_OK) st = sc.out(out, RTEN_F32, xv.ndim, xv.shape, &ov, (out->data == nullptr && dense) ? xv.strides : nullptr);
    if (st == RTEN_OK && numel(&xv) > 0) {
        bool same_layout = dense;
        for (int i = 0; i < xv.ndim && same_layout; i++)
            if (xv.shape[i] != 1 && xv.strides[i] != ov.strides[i]) same_layout = false;
        if (same_layout) {
            st = launch_unary(ctx, op, (const float*)xv.data, (float*)ov.data, numel(&xv));
        } else {
            rten_tensor xc;
            st = sc.contiguous(&xv, &xc);
            if (st == RTEN_OK && is_contiguous(&ov)) {
                st = launch_unary(ctx, op, (const float*)xc.data, (float*)ov.data, numel(&xv));
            } else if (st == RTEN_OK) {
                void* t = nullptr;
                st = temp_alloc(ctx, (size_t)numel(&xv) * 4, &t);
                if (st == RTEN_OK) st = launch_unary(ctx, op, (const float*)xc.data, (float*)t, numel(&xv));
                if (st == RTEN_OK) {
                    long long shape[RTEN_MAX_DIMS], ss[RTEN_MAX_DIMS], ds[RTEN_MAX_DIMS];
                    for (int i = 0; i < xv.ndim; i++) {
                        shape[i] = xc.shape[i];
                        ss[i] = xc.strides[i];
                        ds[i] = ov.strides[i];
                    }
                    st = launch_nd_copy(ctx, 4, t, ov.data, xv.ndim, shape, ss, ds);
                }
            }
        }
    }
    return sc.finish(st);
}

rten_status rten_b200_erf(rten_ctx* ctx, const rten_tensor* x, rten_tensor* out) { return unary_op(ctx, UNARY_ERF, x, out); }
rten_status rten_b200_gelu(rten_ctx* ctx, const rten_tensor* x, int approximate, rten_tensor* out) {
    return unary_op(ctx, approximate ? UNARY_APPROX_GELU : UNARY_GELU, x, out);
}
rten_status rten_b200_relu(rten_ctx* ctx, const rten_tensor* x, rten_tensor* out) { return unary_op(ctx, UNARY_RELU, x, out); }

// ---- Add ----------------------------------------------------------------------------------------------
// Add / Mul with numpy broadcasting (src/ops/binary_elementwise.rs); flags: 0 = Add, 2 = Mul
static rten_status binary_f32(rten_ctx* ctx, const rten_tensor* a, const rten_tensor* b, rten_tensor* out, int flags) {
    RTB_TRY(check_ctx(ctx));
    if (!a || !b || !out) return fail(ctx, RTEN_ERR_MISSING_INPUTS, "missing inputs");
    if (a->dtype != RTEN_F32 || b->dtype != RTEN_F32) return fail(ctx, RTEN_ERR_UNSUPPORTED_TYPE, "unsupported type");
    OpScope sc(ctx);
    rten_tensor av, bv, ov;
    rten_status st = sc.in(a, &av);
    if (st == RTEN_OK) st = sc.in(b, &bv);
    int nd = std::max(a->ndim, b->ndim);
    int64_t shape[RTEN_MAX_DIMS];
    long long sa[RTEN_MAX_DIMS], sb[RTEN_MAX_DIMS];
    for (int i = 0; i < nd && st == RTEN_OK; i++) {
        const int ia = i - (nd - av.ndim), ib = i - (nd - bv.ndim);
        const int64_t da = ia >= 0 ? av.shape[ia] : 1, db = ib >= 0 ? bv.shape[ib] : 1;
        if (da != db && da != 1 && db != 1) st = fail(ctx, RTEN_ERR_INCOMPATIBLE_SHAPES, "Cannot broadcast inputs");
        shape[i] = (da == 0 || db == 0) ? 0 : std::max(da, db);
        sa[i] = (ia >= 0 && da != 1) ? av.strides[ia] : 0;
        sb[i] = (ib >= 0 && db != 1) ? bv.strides[ib] : 0;
    }
    // same-shape dense operands: keep a's layout for the output
    bool same = st == RTEN_OK && av.ndim == bv.ndim && span_elems(&av) == numel(&av);
    for (int i = 0; i < nd && same; i++)
        if (av.shape[i] != bv.shape[i] || (av.shape[i] != 1 && av.strides[i] != bv.strides[i])) same = false;
    if (st == RTEN_OK) st = sc.out(out, RTEN_F32, nd, shape, &ov, (out->data == nullptr && same) ? av.strides : nullptr);
    if (st == RTEN_OK && numel(&ov) > 0) {
        bool flat = same;
        for (int i = 0; i < nd && flat; i++)
            if (ov.shape[i] != 1 && ov.strides[i] != av.strides[i]) flat = false;
        if (flat) {
            st = launch_add_flat(ctx, (const float*)av.data, (const float*)bv.data, (float*)ov.data, numel(&ov), flags);
        } else {
            long long shp[RTEN_MAX_DIMS], sd[RTEN_MAX_DIMS];
            for (int i = 0; i < nd; i++) {
                shp[i] = shape[i];
                sd[i] = ov.strides[i];
            }
            st = launch_nd_add(ctx, (const float*)av.data, (const float*)bv.data, (float*)ov.data, nd, shp, sa, sb, sd, flags);
        }
    }
    return sc.finish(st);
}

rten_status rten_b200_add(rten_ctx* ctx, const rten_tensor* a, const rten_tensor* b, rten_tensor* out) {
    return binary_f32(ctx, a, b, out, 0);
}
rten_status rten_b200_mul(rten_ctx* ctx, const rten_tensor* a, const rten_tensor* b, rten_tensor* out) {
    return binary_f32(ctx, a, b, out, 2);
}

// ---- DynamicQuantizeLinear ---------------------------------------------------------------------------
rten_status rten_b200_range_reset(rten_ctx* ctx, rten_tensor* ranges) {
    RTB_TRY(check_ctx(ctx));
    if (!ranges) return fail(ctx, RTEN_ERR_MISSING_INPUTS, "missing inputs");
    if (ranges->dtype != RTEN_I32 || ranges->device < 0 || !is_contiguous(ranges) || numel(ranges) % 2)
        return fail(ctx, RTEN_ERR_INVALID_VALUE, "ranges must be a contiguous device-resident i32[n, 2]");
    cudaSetDevice(ctx->device);
    return launch_range_reset(ctx, (int*)ranges->data, (int)(numel(ranges) / 2));
}

rten_status rten_b200_dynamic_quantize_linear(rten_ctx* ctx, const rten_tensor* x, rten_tensor* y, rten_tensor* scale,
                                              rten_tensor* zero_point, void* nccl_comm) {
    return rten_b200_dynamic_quantize_linear_ranged(ctx, x, nullptr, y, scale, zero_point, nccl_comm);
}

rten_status rten_b200_dynamic_quantize_linear_ranged(rten_ctx* ctx, const rten_tensor* x, const rten_tensor* range,
                                                     rten_tensor* y, rten_tensor* scale, rten_tensor* zero_point,
                                                     void* nccl_comm) {
    RTB_TRY(check_ctx(ctx));
    if (!x || !y || !scale || !zero_point) return fail(ctx, RTEN_ERR_MISSING_INPUTS, "missing inputs");
    if (x->dtype != RTEN_F32) return fail(ctx, RTEN_ERR_UNSUPPORTED_TYPE, "unsupported type");
    if (range && (range->dtype != RTEN_I32 || numel(range) != 2 || range->device < 0 || !is_contiguous(range)))
        return fail(ctx, RTEN_ERR_INVALID_VALUE, "the range must be a device-resident i32[2]");
    OpScope sc(ctx);
    rten_tensor xv, xc, yv, sv, zv;
    rten_status st = sc.in(x, &xv);
    // The op is elementwise plus an order-independent min / max: a DENSE input in any dim order (e.g. channels-last
    // activations) is processed in memory order and the quantised output keeps the input's strides.
    bool dense = st == RTEN_OK && span_elems(&xv) == numel(&xv) && y->data == nullptr;
    for (int i = 0; i < xv.ndim && dense; i++)
        if (xv.strides[i] <= 0 && xv.shape[i] > 1) dense = false;
    const bool x_cl_dense = st == RTEN_OK && xv.ndim == 4 && xv.strides[1] == 1 && xv.strides[3] == xv.shape[1] &&
                            xv.strides[2] == xv.shape[3] * xv.shape[1] && xv.strides[0] == xv.shape[2] * xv.shape[3] * xv.shape[1];
    if (dense || (x_cl_dense && y->data && y->ndim == 4 && y->strides[1] == 1)) {
        xc = xv;
    } else if (st == RTEN_OK) {
        st = sc.contiguous(&xv, &xc);
    }
    // A caller-provided channels-last output whose rows (b, h) sit at arbitrary pitches -- the interior of a spatially
    // pre-padded buffer, so that the consuming ConvInteger needs no padded copy -- is written row by row.
    bool rows_out = false;
    if (st == RTEN_OK && y->data && xv.ndim == 4 && y->ndim == 4 && y->device >= 0) {
        const int64_t Cc = xv.shape[1], Hh = xv.shape[2], Ww = xv.shape[3];
        rows_out = xv.strides[1] == 1 && xv.strides[3] == Cc && xv.strides[2] == Ww * Cc && xv.strides[0] == Hh * Ww * Cc &&
                   y->strides[1] == 1 && y->strides[3] == Cc && !is_contiguous(y) &&
                   !(y->strides[2] == Ww * Cc && y->strides[0] == Hh * Ww * Cc);
        if (rows_out) xc = xv;
    }
    if (st == RTEN_OK) st = sc.out(y, RTEN_U8, xv.ndim, xv.shape, &yv, dense ? xv.strides : nullptr);
    if (st == RTEN_OK) st = sc.out(scale, RTEN_F32, 0, nullptr, &sv, nullptr);
    if (st == RTEN_OK) st = sc.out(zero_point, RTEN_U8, 0, nullptr, &zv, nullptr);
    if (st == RTEN_OK && !dense && !rows_out && !is_contiguous(&yv)) st = fail(ctx, RTEN_ERR_UNSUPPORTED_OUTPUT, "quantized output must be contiguous");
    if (st == RTEN_OK) {
        const long long n = numel(&xv);
        if (n == 0) {
            // quantize.rs:378-386: scale 1, zero point 0
            const float one = 1.0f;
            RTB_CUDA(ctx, cudaMemcpyAsync(sv.data, &one, 4, cudaMemcpyHostToDevice, rtb::launch_stream(ctx)));
            RTB_CUDA(ctx, cudaMemsetAsync(zv.data, 0, 1, rtb::launch_stream(ctx)));
        } else if (!nccl_comm && !range && !rows_out && n <= 16384) {
            st = launch_dql_small(ctx, (const float*)xc.data, (uint8_t*)yv.data, (int)n, (float*)sv.data, (uint8_t*)zv.data);
        } else {
            // `range`: the producer of x already accumulated (min, max) in its epilogue -- no pass over x for it
            int* mm = range ? (int*)range->data : nullptr;
            if (!mm) {
                st = temp_alloc(ctx, 8, (void**)&mm);
                if (st == RTEN_OK) st = launch_minmax(ctx, (const float*)xc.data, n, mm);
            }
            // batch-sharded run: the range is the range of the whole (unsharded) tensor
            if (st == RTEN_OK && nccl_comm) st = comm_allreduce_minmax(ctx, reinterpret_cast<rten_comm*>(nccl_comm), mm);
            if (st == RTEN_OK && rows_out)
                st = launch_dql_quantize_rows(ctx, (const float*)xc.data, (uint8_t*)yv.data, xv.shape[0] * xv.shape[2],
                                              (int)(xv.shape[3] * xv.shape[1]), (int)xv.shape[2], yv.strides[2], yv.strides[0], mm,
                                              (float*)sv.data, (uint8_t*)zv.data);
            else if (st == RTEN_OK)
                st = launch_dql_quantize(ctx, (const float*)xc.data, (uint8_t*)yv.data, n, mm, (float*)sv.data, (uint8_t*)zv.data);
        }
    }
    return sc.finish(st);
}

// ---- pooling / gather ---------------------------------------------------------------------------------
rten_status rten_b200_max_pool(rten_ctx* ctx, const rten_tensor* x, const int32_t kernel[2], const int32_t pads[4],
                               const int32_t strides[2], rten_tensor* out) {
    RTB_TRY(check_ctx(ctx));
    if (!x || !out) return fail(ctx, RTEN_ERR_MISSING_INPUTS, "missing inputs");
    if (x->dtype != RTEN_F32) return fail(ctx, RTEN_ERR_UNSUPPORTED_TYPE, "unsupported type");
    if (x->ndim != 4) return fail(ctx, RTEN_ERR_INVALID_VALUE, "input must have 4 dims (NCHW)");
    OpScope sc(ctx);
    rten_tensor xv, ov;
    rten_status st = sc.in(x, &xv);
    int64_t OH = 0, OW = 0, p0, p1;
    if (st == RTEN_OK) st = axis_out(ctx, xv.shape[2], kernel[0], strides[0], false, pads[0], pads[2], 1, &OH, &p0, &p1);
    if (st == RTEN_OK) st = axis_out(ctx, xv.shape[3], kernel[1], strides[1], false, pads[1], pads[3], 1, &OW, &p0, &p1);
    if (st == RTEN_OK) {
        const int64_t B = xv.shape[0], C = xv.shape[1];
        int64_t oshape[4] = {B, C, OH, OW};
        const bool cl = xv.strides[1] == 1 && C > 1;
        int64_t pref[4] = {C * OH * OW, OH * OW, OW, 1};
        if (cl) {
            pref[0] = OH * OW * C;
            pref[1] = 1;
            pref[2] = OW * C;
            pref[3] = C;
        }
        st = sc.out(out, RTEN_F32, 4, oshape, &ov, out->data ? nullptr : pref);
        if (st == RTEN_OK) {
            PoolParams p;
            p.B = (int)B;
            p.C = (int)C;
            p.H = (int)xv.shape[2];
            p.W = (int)xv.shape[3];
            p.OH = (int)OH;
            p.OW = (int)OW;
            p.kh = kernel[0];
            p.kw = kernel[1];
            p.sy = strides[0];
            p.sx = strides[1];
            p.pt = pads[0];
            p.pl = pads[1];
            p.xs_b = xv.strides[0];
            p.xs_c = xv.strides[1];
            p.xs_h = xv.strides[2];
            p.xs_w = xv.strides[3];
            p.ys_b = ov.strides[0];
            p.ys_c = ov.strides[1];
            p.ys_h = ov.strides[2];
            p.ys_w = ov.strides[3];
            p.channels_fastest = ov.strides[1] == 1;
            st = launch_maxpool(ctx, (const float*)xv.data, (float*)ov.data, p);
        }
    }
    return sc.finish(st);
}

rten_status rten_b200_global_average_pool(rten_ctx* ctx, const rten_tensor* x, rten_tensor* out) {
    RTB_TRY(check_ctx(ctx));
    if (!x || !out) return fail(ctx, RTEN_ERR_MISSING_INPUTS, "missing inputs");
    if (x->dtype != RTEN_F32) return fail(ctx, RTEN_ERR_UNSUPPORTED_TYPE, "unsupported type");
    if (x->ndim != 4) return fail(ctx, RTEN_ERR_INVALID_VALUE, "input must have 4 dims (NCHW)");
    OpScope sc(ctx);
    rten_tensor xv, ov;
    rten_status st = sc.in(x, &xv);
    if (st == RTEN_OK) {
        const int64_t B = xv.shape[0], C = xv.shape[1], H = xv.shape[2], W = xv.shape[3];
        int64_t oshape[4] = {B, C, 1, 1};
        st = sc.out(out, RTEN_F32, 4, oshape, &ov, nullptr);
        if (st == RTEN_OK && !is_contiguous(&ov)) st = fail(ctx, RTEN_ERR_UNSUPPORTED_OUTPUT, "pooled output must be contiguous");
        if (st == RTEN_OK && B * C > 0) {
            rten_tensor src = xv;
            if (!(xv.strides[2] == W * xv.strides[3])) {  // need a uniform stride over (h, w)
                st = sc.contiguous(&xv, &src);
            }
            if (st == RTEN_OK)
                st = launch_row_mean(ctx, (const float*)src.data, (float*)ov.data, B * C, (int)(H * W), C, src.strides[0],
                                     src.strides[1], src.strides[3]);
        }
    }
    return sc.finish(st);
}

rten_status rten_b200_gather_rows(rten_ctx* ctx, const rten_tensor* table, const rten_tensor* idx, rten_tensor* out) {
    RTB_TRY(check_ctx(ctx));
    if (!table || !idx || !out) return fail(ctx, RTEN_ERR_MISSING_INPUTS, "missing inputs");
    if (table->dtype != RTEN_F32 || idx->dtype != RTEN_I32) return fail(ctx, RTEN_ERR_UNSUPPORTED_TYPE, "unsupported type");
    if (table->ndim != 2) return fail(ctx, RTEN_ERR_INVALID_VALUE, "gather_rows expects a 2-D table");
    OpScope sc(ctx);
    rten_tensor tv, iv, ic, ov;
    rten_status st = sc.in(table, &tv);
    if (st == RTEN_OK) st = sc.in(idx, &iv);
    if (st == RTEN_OK) st = sc.contiguous(&iv, &ic);
    if (st == RTEN_OK) {
        int64_t oshape[RTEN_MAX_DIMS];
        if (iv.ndim + 1 > RTEN_MAX_DIMS) return sc.finish(fail(ctx, RTEN_ERR_INVALID_VALUE, "tensor rank out of range"));
        for (int i = 0; i < iv.ndim; i++) oshape[i] = iv.shape[i];
        oshape[iv.ndim] = tv.shape[1];
        st = sc.out(out, RTEN_F32, iv.ndim + 1, oshape, &ov, nullptr);
        if (st == RTEN_OK && !is_contiguous(&ov)) st = fail(ctx, RTEN_ERR_UNSUPPORTED_OUTPUT, "gather output must be contiguous");
        if (st == RTEN_OK)
            st = launch_gather_rows(ctx, (const float*)tv.data, (const int*)ic.data, (float*)ov.data, numel(&iv),
                                    (int)tv.shape[1], tv.strides[0], tv.strides[1], tv.shape[0]);
    }
    return sc.finish(st);
}

rten_status rten_b200_scatter_rows(rten_ctx* ctx, rten_tensor* table, const rten_tensor* idx, const rten_tensor* src) {
    RTB_TRY(check_ctx(ctx));
    if (!table || !idx || !src) return fail(ctx, RTEN_ERR_MISSING_INPUTS, "missing inputs");
    if (table->dtype != RTEN_F32 || src->dtype != RTEN_F32 || idx->dtype != RTEN_I32)
        return fail(ctx, RTEN_ERR_UNSUPPORTED_TYPE, "unsupported type");
    if (table->ndim != 2 || src->ndim != 2 || idx->ndim != 1) return fail(ctx, RTEN_ERR_INVALID_VALUE, "scatter_rows expects 2-D table / updates and 1-D indices");
    if (src->shape[0] != idx->shape[0] || src->shape[1] != table->shape[1])
        return fail(ctx, RTEN_ERR_INCOMPATIBLE_SHAPES, "updates do not match the indices / table width");
    if (table->device < 0) return fail(ctx, RTEN_ERR_UNSUPPORTED_OUTPUT, "the table must be device resident (updated in place)");
    OpScope sc(ctx);
    rten_tensor iv, ic, sv;
    rten_status st = sc.in(idx, &iv);
    if (st == RTEN_OK) st = sc.contiguous(&iv, &ic);
    if (st == RTEN_OK) st = sc.in(src, &sv);
    if (st == RTEN_OK)
        st = launch_scatter_rows(ctx, (float*)table->data, (const int*)ic.data, (const float*)sv.data, iv.shape[0],
                                 (int)table->shape[1], table->strides[0], table->strides[1], sv.strides[0], sv.strides[1],
                                 table->shape[0]);
    return sc.finish(st);
}

}  // extern "C"
