// Launchers of the HBM-bound kernels (rowops.cu).  All pointers are device pointers; all launches
// go to ctx->stream.
#pragma once
#include <cstdint>

#include "common.h"

namespace rtb {

enum { UNARY_ERF = 0, UNARY_GELU = 1, UNARY_APPROX_GELU = 2, UNARY_RELU = 3 };

// Softmax over the last (contiguous) axis of x viewed as [rows, n]; optional mask broadcast over
// up to 4 leading dims: row index is decomposed over lead[0..nlead) (row-major), mask element =
// mask[sum idx_d * mstride[d] + i * mstride_last].
rten_status launch_softmax(rten_ctx* ctx, const float* x, float* y, long long rows, int n, int flush_nan,
                           const float* mask, int nlead, const long long* lead, const long long* mstride,
                           long long mstride_last);
// gamma_sp / beta_sp: scalar scale / bias resident on the device (then gamma / beta are null and the host scalars unused)
rten_status launch_layer_norm(rten_ctx* ctx, const float* x, float* y, long long rows, int n, const float* gamma,
                              float gamma_scalar, const float* beta, float beta_scalar, float eps,
                              const float* gamma_sp = nullptr, const float* beta_sp = nullptr);
// y[row] = Sum(x_row) / n in the reference's Sum order; row r starts at
// x + (r / rows_inner) * s_outer + (r % rows_inner) * s_inner, elements kstride apart.
rten_status launch_row_mean(rten_ctx* ctx, const float* x, float* y, long long rows, int n, long long rows_inner,
                            long long s_outer, long long s_inner, long long kstride);
rten_status launch_unary(rten_ctx* ctx, int op, const float* x, float* y, long long n);
rten_status launch_nd_copy(rten_ctx* ctx, int esize, const void* src, void* dst, int ndim, const long long* shape,
                           const long long* sstride, const long long* dstride);
rten_status launch_nd_add(rten_ctx* ctx, const float* a, const float* b, float* d, int ndim, const long long* shape,
                          const long long* sa, const long long* sb, const long long* sd, int relu);
rten_status launch_add_flat(rten_ctx* ctx, const float* a, const float* b, float* d, long long n, int relu);
rten_status launch_minmax(rten_ctx* ctx, const float* x, long long n, int* mm /* 2 ordered ints */);
// `xch` (batch-sharded runs): the kernel first exchanges the local range in `mm` with the other ranks (comm_device.cuh)
struct RangeExchange;
rten_status launch_dql_quantize(rten_ctx* ctx, const float* x, uint8_t* y, long long n, int* mm,
                                float* scale_out, uint8_t* zp_out, const RangeExchange* xch = nullptr);
rten_status launch_rowsum8(rten_ctx* ctx, const void* a, int is_signed, long long rows, int K, long long ld, int* out);
rten_status launch_zp_to_i32(rten_ctx* ctx, const void* zp, int is_signed, int n, long long zs, int* out);
rten_status launch_fill8(rten_ctx* ctx, void* p, long long n, uint8_t v);
rten_status launch_cast_scale(rten_ctx* ctx, const int* in, float* out, long long n, int cols, const float* scale,
                              int scale_len);

struct Im2ColParams {
    int B, C, H, W, OH, OW, kh, kw, sy, sx, dy, dx, pt, pl;
    int c0;    // first input channel (group offset)
    int kpad;  // output row pitch in elements (>= kh*kw*C, zero filled beyond)
    long long xs_b, xs_c, xs_h, xs_w;  // input element strides
};
rten_status launch_im2col(rten_ctx* ctx, int esize, const void* x, void* out, const Im2ColParams& p, int pad_value);

rten_status launch_smallc_pad(rten_ctx* ctx, const float* x, float* xp, int B, int C, int H, int W, int Wp, int pl,
                              long long xs_b, long long xs_c, long long xs_h, long long xs_w);
rten_status launch_smallc_pack_w(rten_ctx* ctx, const float* w, float* wp, int O, int C, int kh, int kw, long long ws_o,
                                 long long ws_c, long long ws_h, long long ws_w);

struct PoolParams {
    int B, C, H, W, OH, OW, kh, kw, sy, sx, pt, pl;
    long long xs_b, xs_c, xs_h, xs_w, ys_b, ys_c, ys_h, ys_w;
    int channels_fastest;  // thread -> element mapping that keeps warps coalesced for NHWC memory
};
rten_status launch_maxpool(rten_ctx* ctx, const float* x, float* y, const PoolParams& p);
rten_status launch_gather_rows(rten_ctx* ctx, const float* table, const int* idx, float* out, long long nidx,
                               int width, long long t_rs, long long t_cs, long long rows);

// 3xTF32 operand split: dst contiguous [d3][d2][d1][3 * d0p]; role 0 = [lo|hi|hi], 1 = [hi|lo|hi]
rten_status launch_tf32x3_split(rten_ctx* ctx, const float* x, float* y, const long long dims[4], const long long strides[4],
                                long long d0p, int role);

rten_status launch_dql_small(rten_ctx* ctx, const float* x, uint8_t* y, int n, float* scale_out, uint8_t* zp_out);

rten_status launch_scatter_rows(rten_ctx* ctx, float* table, const int* idx, const float* src, long long nidx, int width,
                                long long t_rs, long long t_cs, long long s_rs, long long s_cs, long long rows);

rten_status launch_range_reset(rten_ctx* ctx, int* mm, int pairs);

rten_status launch_smallc8_pad(rten_ctx* ctx, const void* x, void* xp, int B, int C, int H, int W, int Hp, int Wp, int pt,
                               int pl, long long xs_b, long long xs_c, long long xs_h, long long xs_w, int pad_value);
rten_status launch_smallc8_pack_w(rten_ctx* ctx, const void* w, void* wp, int O, int C, int kh, int kw, long long ws_o,
                                  long long ws_c, long long ws_h, long long ws_w);

rten_status launch_dql_quantize_rows(rten_ctx* ctx, const float* x, uint8_t* y, long long rows, int row_len, int rows_inner,
                                     long long y_inner, long long y_outer, int* mm, float* scale_out, uint8_t* zp_out,
                                     const RangeExchange* xch = nullptr);

}  // namespace rtb
