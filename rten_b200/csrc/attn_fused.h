// Launcher of the fused encoder attention kernel (attn_fused.cu): softmax(scale * Q K^T + mask) V for 128 keys of head
// size 64, queries in tiles of 128, single-pass TF32 products with the reference's softmax arithmetic in between.
#pragma once
#include "umma_gemm.h"

namespace rtb {

struct AttnFusedLaunch {
    int B = 0, heads = 0, q_seq = 0, kv_seq = 0, dh = 0;
    OperandDesc q;   // (d, s, h, b): head dimension contiguous
    OperandDesc k;   // (d, s, h, b)
    OperandDesc vt;  // (s, d, h, b): KEY dimension contiguous (the value tensor stored transposed) -- used when v == null
    const float* v = nullptr;  // natural value tensor, head dimension contiguous, element strides v_b / v_h / v_s
    long long v_b = 0, v_h = 0, v_s = 0;
    const float* mask = nullptr;  // additive, [B, kv_seq] with row stride m_b (0 = one row for every batch), or null
    long long m_b = 0;
    float scale = 1.0f;
    float* out = nullptr;  // element strides o_b, o_h, o_s; head dimension contiguous
    long long o_b = 0, o_h = 0, o_s = 0;
};
bool attn_fused_supported(const AttnFusedLaunch& L);
rten_status launch_attn_fused(rten_ctx* ctx, const AttnFusedLaunch& L);

}  // namespace rtb
