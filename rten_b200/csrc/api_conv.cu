// Operator entry points of the C ABI, Conv family (Conv, ConvInteger(ToFloat), weight prepack) and pooling:
// shape / argument validation with the reference's error strings,
// operand normalisation (K-major, TMA-addressable), kernel dispatch.  Mirrors, per function, the
// reference operator named in include/rten_b200.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "api_shared.h"
#include "api_util.h"
#include "rowops.h"
#include "skinny.h"
#include "umma_gemm.h"

using namespace rtb;
using namespace rtb::api;

namespace {

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
    const bool smallc_ok = !implicit_ok && A.kind == 0 && groups == 1 && Cg <= 4 && kw * 4 <= 32 && dil[1] == 1 &&
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
        if (ctx->f32_mode == RTEN_F32_TF32X3 && !getenv("RTEN_B200_X3_THREE_PLANES")) {
            // 3xTF32: the low parts of the padded copy, once (the window view below overlaps itself 8 / stride times)
            float* xlo = nullptr;
            const long long n = (long long)B * H * Wp * 4;
            RTB_TRY(temp_alloc(ctx, (size_t)n * 4, (void**)&xlo));
            const long long fd[4] = {n, 1, 1, 1}, fs[4] = {1, n, n, n};
            RTB_TRY(launch_tf32x3_split(ctx, xp, xlo, fd, fs, n, 2));
            L.a_lo_base = xlo;
        }
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
            if (A.pw && A.kind == 0 && groups == 1 && wg == A.pw->data) L.b_x3_slot = &const_cast<rten_packed*>(A.pw)->x3;
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

extern "C" {

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

}  // extern "C"
