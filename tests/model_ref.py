"""The same post-fusion op lists as rten_b200/graphs.py, executed by the CPU oracle (NCHW, separate
Conv / Add / Relu ops, the reference's own arithmetic).  Used for whole-model parity and as the
bench's cpu_baseline / reference arm."""
import math

import numpy as np


def resnet50_oracle(oracle, spec, x, arena=None):
    """`arena` (oracle.Arena): recycle operator outputs from the previous pass and run Relu / Add in place, as the
    reference's BufferPool and in-place operators do (same arithmetic)."""
    n = [0]

    def conv(c, t):
        n[0] += 1
        out = None
        if arena is not None:
            B, _, H, W = t.shape
            O, _, kh, kw = c.w.shape
            oh = (H + 2 * c.pad - kh) // c.stride + 1
            ow = (W + 2 * c.pad - kw) // c.stride + 1
            out = arena.get(("conv", n[0]), (B, O, oh, ow))
        return oracle.conv(t, c.w, c.b, [c.pad] * 4, 1, (c.stride, c.stride), (1, 1), out=out)

    def relu(t):
        return oracle.relu(t, out=t if arena is not None else None)

    def add(a, b):
        return oracle.add(a, b, out=a if arena is not None else None)

    y = relu(conv(spec.stem, x))
    y = oracle.max_pool(y, (3, 3), [1, 1, 1, 1], (2, 2))
    for b in spec.blocks:
        ident = y if b.down is None else conv(b.down, y)
        t = relu(conv(b.c1, y))
        t = relu(conv(b.c2, t))
        y = relu(add(conv(b.c3, t), ident))
    p = oracle.global_average_pool(y)
    return oracle.gemm_op(p.reshape(p.shape[0], p.shape[1]), spec.fc_w, spec.fc_b, 1.0, 1.0, False, True)


def mnist_oracle(oracle, w, x):
    """configs[0]: the reference's MNIST CNN (rten-onnx/test-data/mnist.onnx) with the reference's operators."""
    y = oracle.relu(oracle.conv(x, w["conv1.weight"], w["conv1.bias"], [1, 1, 1, 1], 1, (1, 1), (1, 1)))
    y = oracle.max_pool(y, (2, 2), [0, 0, 0, 0], (2, 2))
    y = oracle.relu(oracle.conv(y, w["conv2.weight"], w["conv2.bias"], [1, 1, 1, 1], 1, (1, 1), (1, 1)))
    y = oracle.max_pool(y, (2, 2), [0, 0, 0, 0], (2, 2))
    y = oracle.relu(oracle.conv(y, w["pw.weight"], w["pw.bias"], [0, 0, 0, 0], 1, (1, 1), (1, 1)))
    p = oracle.global_average_pool(y)
    return oracle.gemm_op(p.reshape(p.shape[0], p.shape[1]), w["fc.weight"], w["fc.bias"], 1.0, 1.0, False, True)


def resnet50_int8_oracle(oracle, qspec, x, w_zero_points=True):
    """configs[3] (rten_b200/graphs.py ResNet50Int8Runner) with the reference's operators, unfused, NCHW.
    -> (logits, pooled features)."""
    f32 = np.float32

    def conv(c, t, relu, residual=None):
        tq, ts, tz = oracle.dynamic_quantize_linear(t)
        scale = f32(ts) * f32(c.w_scale)
        wz = np.zeros(c.wq.shape[0], np.int8) if w_zero_points else None
        y = oracle.conv_integer_to_float(tq, c.wq, tz, wz, scale, padding=[c.pad] * 4, groups=1,
                                         strides=(c.stride, c.stride), dilations=(1, 1))
        y = oracle.add(y, c.b.reshape(1, -1, 1, 1))
        if residual is not None:
            y = oracle.add(y, residual)
        return oracle.relu(y) if relu else y

    y = conv(qspec.stem, x, True)
    y = oracle.max_pool(y, (3, 3), [1, 1, 1, 1], (2, 2))
    for b in qspec.blocks:
        ident = y if b.down is None else conv(b.down, y, False)
        t = conv(b.c1, y, True)
        t = conv(b.c2, t, True)
        y = conv(b.c3, t, True, residual=ident)
    p = oracle.global_average_pool(y)
    p = p.reshape(p.shape[0], p.shape[1])
    return oracle.gemm_op(p, qspec.fc_w, qspec.fc_b, 1.0, 1.0, False, True), p


def bert_oracle(oracle, spec, input_ids, token_type_ids, add_mask):
    B, S = input_ids.shape
    H, nh = spec.hidden, spec.heads
    dh = H // nh
    x = oracle.add(oracle.add(spec.word_emb[input_ids], spec.pos_emb[:S]), spec.type_emb[token_type_ids])
    x = oracle.layer_norm(x, spec.emb_g, spec.emb_b, -1, spec.eps).reshape(B * S, H)
    scale = 1.0 / math.sqrt(dh)
    for L in spec.layers:
        q = oracle.matmul(x, L.wq, L.bq).reshape(B, S, nh, dh).transpose(0, 2, 1, 3)
        k = oracle.matmul(x, L.wk, L.bk).reshape(B, S, nh, dh).transpose(0, 2, 3, 1)
        v = oracle.matmul(x, L.wv, L.bv).reshape(B, S, nh, dh).transpose(0, 2, 1, 3)
        probs = oracle.add_softmax(oracle.matmul(q, k, None, scale), add_mask)
        att = oracle.matmul(probs, v).transpose(0, 2, 1, 3).reshape(B * S, H)
        x = oracle.layer_norm(oracle.add(oracle.matmul(att, L.wo, L.bo), x), L.ln1_g, L.ln1_b, -1, spec.eps)
        h = oracle.gelu(oracle.matmul(x, L.w1, L.b1))
        x = oracle.layer_norm(oracle.add(oracle.matmul(h, L.w2, L.b2), x), L.ln2_g, L.ln2_b, -1, spec.eps)
    return x.reshape(B, S, H)


class gpt2_int8_decoder:
    """configs[4] with the reference's operators, stateful: the constructor runs the prompt (prefill), `step(ids)` feeds
    the next [B,T] token block against the `past_key_values` kept so far.  `.logits` = last-position logits [B,vocab]."""

    def __init__(self, oracle, spec, prompt_ids):
        self.oracle, self.spec = oracle, spec
        self.past = [None] * len(spec.layers)
        self.P = 0
        self.logits = self.step(prompt_ids)

    def _linear(self, x, l, gelu=False, residual=None):
        oracle, f32 = self.oracle, np.float32
        xq, xs, xz = oracle.dynamic_quantize_linear(x)
        y = oracle.matmul_integer_to_float(xq, l.wq, xz, None, (f32(xs) * l.w_scale).astype(f32))
        if l.b is not None:
            y = oracle.add(y, l.b)
        if residual is not None:
            y = oracle.add(y, residual)
        return oracle.gelu(y, True) if gelu else y

    def step(self, ids):
        oracle, spec, f32 = self.oracle, self.spec, np.float32
        H, nh = spec.hidden, spec.heads
        dh = H // nh
        B, T = ids.shape
        P = self.P
        L = P + T
        x = oracle.add(spec.wte[ids], spec.wpe[P:L]).reshape(B * T, H)
        mask = np.where(np.arange(L)[None, :] <= (P + np.arange(T))[:, None], 0.0, -np.inf).astype(f32).reshape(1, 1, T, L)
        for li, ly in enumerate(spec.layers):
            h = oracle.layer_norm(x, ly.ln1_g, ly.ln1_b, -1, spec.eps)
            qkv = self._linear(h, ly.attn).reshape(B, T, 3, nh, dh)
            q, k, v = (qkv[:, :, i].transpose(0, 2, 1, 3) for i in range(3))
            if self.past[li] is not None:
                k = np.concatenate([self.past[li][0], k], 2)
                v = np.concatenate([self.past[li][1], v], 2)
            self.past[li] = (k, v)
            probs = oracle.add_softmax(oracle.matmul(q, k.transpose(0, 1, 3, 2), None, 1.0 / math.sqrt(dh)), mask)
            att = oracle.matmul(probs, v).transpose(0, 2, 1, 3).reshape(B * T, H)
            x = self._linear(att, ly.proj, residual=x)
            h = oracle.layer_norm(x, ly.ln2_g, ly.ln2_b, -1, spec.eps)
            x = self._linear(self._linear(h, ly.fc, gelu=True), ly.fc2, residual=x)
        last = oracle.layer_norm(x.reshape(B, T, H)[:, -1], spec.lnf_g, spec.lnf_b, -1, spec.eps)
        self.P = L
        self.logits = self._linear(last, spec.lm_head)
        return self.logits


def gpt2_int8_oracle(oracle, spec, steps):
    """-> list of last-position logits [B,vocab], one per token block of `steps` (prefill, then decode steps)."""
    dec = gpt2_int8_decoder(oracle, spec, steps[0])
    outs = [dec.logits]
    for ids in steps[1:]:
        outs.append(dec.step(ids))
    return outs
