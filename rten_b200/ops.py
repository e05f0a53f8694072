"""Host-side mirror of RTen's operator interface for the hot path, over the C ABI.

Each class below corresponds to the reference operator of the same name and keeps its attribute
names (src/ops/matmul.rs, conv.rs, norm.rs, attention.rs, unary_elementwise.rs, quantize.rs):
`op.run(ctx, *inputs) -> outputs` plays `Operator::run(&OpRunContext)` (src/operator.rs:498),
`op.prepack(ctx, index, input)` plays `Operator::prepack` (:587-601), and failures raise `OpError`
carrying the reference's `OpError` variant and message.

Inputs may be numpy arrays (host tensors: the library stages them through HBM inside the call)
or `DeviceTensor`s (resident in HBM).  Outputs are `DeviceTensor`s allocated from the context's
pool; `.numpy()` copies them back.  Nothing here computes: no CUDA library => exception.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import RTEN_DEVICE_HOST, RTEN_F32, RTEN_I8, RTEN_I32, RTEN_U8, RtenAttentionParams, RtenConvParams, RtenTensor

_NP2RT = {np.dtype(np.float32): RTEN_F32, np.dtype(np.int32): RTEN_I32, np.dtype(np.int8): RTEN_I8,
          np.dtype(np.uint8): RTEN_U8}
_RT2NP = {v: k for k, v in _NP2RT.items()}

ACT_NONE, ACT_RELU, ACT_GELU, ACT_GELU_TANH = 0, 1, 2, 3


class OpError(Exception):
    """`OpError` (src/operator.rs:116-144)."""

    def __init__(self, status: int, msg: str):
        self.status = status
        self.kind = _lib.STATUS_NAMES.get(status, str(status))
        self.msg = msg
        super().__init__(f"{self.kind}: {msg}")


class Context:
    """One per host thread / stream (= OpRunContext + BufferPool, src/operator.rs:328-390)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None, workspace_bytes: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        st = self.lib.rten_b200_ctx_create(device, C.c_void_p(stream) if stream else None, workspace_bytes, C.byref(h))
        if st != 0:
            raise OpError(st, "rten_b200_ctx_create failed: a B200 (sm_100a) with a working driver is required")
        self.handle = h
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            self.lib.rten_b200_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, st: int):
        if st != 0:
            raise OpError(st, self.lib.rten_b200_last_error(self.handle).decode())

    def sync(self):
        self.check(self.lib.rten_b200_sync(self.handle))

    def set_f32_mode(self, tf32x3: bool):
        """True (the library default): 3xTF32 error-compensated products, f32-grade accuracy at 1/3 of the tensor rate.
        False: single-pass TF32 -- an explicit opt-in to 10-bit operand mantissas."""
        self.check(self.lib.rten_b200_set_f32_mode(self.handle, 1 if tf32x3 else 0))

    def forced_plan_counts(self):
        """(matched, unmatched) launches under the RTEN_B200_FORCE_* sweep knobs."""
        a, b = C.c_uint64(), C.c_uint64()
        self.check(self.lib.rten_b200_debug_forced_plans(self.handle, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def set_autotune(self, enable: bool = True):
        """Time candidate launch plans the first time each MatMul / Conv problem is seen (outside graph capture)."""
        self.check(self.lib.rten_b200_set_autotune(self.handle, 1 if enable else 0))

    def save_plans(self, path: str):
        self.check(self.lib.rten_b200_save_plans(self.handle, path.encode()))

    def load_plans(self, path: str):
        self.check(self.lib.rten_b200_load_plans(self.handle, path.encode()))

    @property
    def launches(self) -> int:
        return int(self.lib.rten_b200_launch_count(self.handle))

    # ---- memory
    def alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self.check(self.lib.rten_b200_alloc(self.handle, nbytes, C.byref(p)))
        return p.value

    def free(self, ptr: int):
        self.check(self.lib.rten_b200_free(self.handle, C.c_void_p(ptr)))

    def pinned_empty(self, shape, dtype) -> np.ndarray:
        """Pinned host array (for asynchronous staging of host tensors)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        self.check(self.lib.rten_b200_host_alloc(self.handle, max(n, 16), C.byref(p)))
        buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        return arr

    def empty(self, shape, dtype=np.float32, strides=None) -> "DeviceTensor":
        dtype = np.dtype(dtype)
        shape = tuple(int(s) for s in shape)
        n = int(np.prod(shape)) if len(shape) else 1
        ptr = self.alloc(max(n, 1) * dtype.itemsize)
        if strides is None:
            strides = _contig(shape)
        return DeviceTensor(self, ptr, shape, tuple(strides), dtype, owner=True)

    def to_device(self, arr, channels_last: bool = False) -> "DeviceTensor":
        arr = np.asarray(arr)
        strides = None
        if channels_last and arr.ndim == 4:
            b, c, h, w = arr.shape
            strides = (h * w * c, 1, w * c, c)
        t = self.empty(arr.shape, arr.dtype, strides)
        t.copy_from(arr)
        return t

    # ---- CUDA graph over an op list
    def graph_begin(self):
        self.check(self.lib.rten_b200_graph_begin(self.handle))

    def graph_end(self) -> "Graph":
        g = C.c_void_p()
        self.check(self.lib.rten_b200_graph_end(self.handle, C.byref(g)))
        return Graph(self, g)


class Graph:
    def __init__(self, ctx: Context, handle):
        self.ctx, self.handle = ctx, handle

    def launch(self):
        self.ctx.check(self.ctx.lib.rten_b200_graph_launch(self.ctx.handle, self.handle))

    def __del__(self):
        try:
            if self.handle and self.ctx.handle:
                self.ctx.lib.rten_b200_graph_destroy(self.handle)
        except Exception:
            pass


def _contig(shape):
    s, out = 1, []
    for d in reversed(shape):
        out.append(s)
        s *= d
    return tuple(reversed(out))


class DeviceTensor:
    """A strided tensor resident in HBM (element strides, like rten-tensor layouts)."""

    def __init__(self, ctx: Context, ptr: int, shape, strides, dtype, owner: bool, base=None):
        self.ctx, self.ptr, self.shape, self.strides, self.dtype = ctx, ptr, tuple(shape), tuple(strides), np.dtype(dtype)
        self.owner = owner
        self.base = base  # keeps the owning tensor alive for views

    def __del__(self):
        try:
            if self.owner and self.ptr and self.ctx.handle:
                self.ctx.lib.rten_b200_free(self.ctx.handle, C.c_void_p(self.ptr))
        except Exception:
            pass

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape)) if self.shape else 1

    def desc(self) -> RtenTensor:
        return _desc(self.ptr, self.dtype, self.shape, self.strides, self.ctx.device)

    def view(self, shape, strides, offset_elems: int = 0) -> "DeviceTensor":
        return DeviceTensor(self.ctx, self.ptr + offset_elems * self.dtype.itemsize, shape, strides, self.dtype, False,
                            base=self.base if self.base is not None else self)

    def permute(self, *axes) -> "DeviceTensor":
        return self.view([self.shape[a] for a in axes], [self.strides[a] for a in axes])

    def reshape(self, *shape) -> "DeviceTensor":
        assert self.is_contiguous(), "reshape needs a contiguous tensor"
        shape = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
        assert int(np.prod(shape)) == self.size
        return self.view(shape, _contig(shape))

    def is_contiguous(self) -> bool:
        return all(s == 1 or st == c for s, st, c in zip(self.shape, self.strides, _contig(self.shape)))

    def copy_from(self, arr: np.ndarray):
        arr = np.ascontiguousarray(arr, dtype=self.dtype).reshape(np.shape(arr))  # (ascontiguousarray turns 0-d into 1-d)
        assert arr.shape == self.shape
        src = _desc(arr.ctypes.data, arr.dtype, arr.shape, _contig(arr.shape), RTEN_DEVICE_HOST)
        dst = self.desc()
        self.ctx.check(self.ctx.lib.rten_b200_copy(self.ctx.handle, C.byref(src), C.byref(dst)))

    def assign(self, src: "DeviceTensor"):
        """Strided device-to-device copy of `src` (same shape) into this view (e.g. appending to a KV cache)."""
        assert tuple(src.shape) == tuple(self.shape) and src.dtype == self.dtype
        a, b = src.desc(), self.desc()
        self.ctx.check(self.ctx.lib.rten_b200_copy(self.ctx.handle, C.byref(a), C.byref(b)))

    def numpy(self) -> np.ndarray:
        out = np.empty(self.shape, self.dtype)
        dst = _desc(out.ctypes.data, out.dtype, out.shape, _contig(out.shape), RTEN_DEVICE_HOST)
        src = self.desc()
        self.ctx.check(self.ctx.lib.rten_b200_copy(self.ctx.handle, C.byref(src), C.byref(dst)))
        return out


def from_torch(ctx: Context, t) -> DeviceTensor:
    """Borrow a CUDA torch tensor's storage (torch is plumbing for device memory only)."""
    import torch

    dt = {torch.float32: np.float32, torch.int32: np.int32, torch.int8: np.int8, torch.uint8: np.uint8}[t.dtype]
    return DeviceTensor(ctx, t.data_ptr(), tuple(t.shape), tuple(t.stride()), dt, owner=False, base=t)


def _desc(ptr, dtype, shape, strides, device) -> RtenTensor:
    d = RtenTensor()
    d.data = ptr
    d.dtype = _NP2RT[np.dtype(dtype)]
    d.ndim = len(shape)
    for i, (s, st) in enumerate(zip(shape, strides)):
        d.shape[i] = int(s)
        d.strides[i] = int(st)
    d.device = device
    return d


class _Args:
    """Keeps numpy inputs alive for the duration of a call and builds descriptors."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.keep = []

    def t(self, x):
        if x is None:
            return None
        if isinstance(x, DeviceTensor):
            d = x.desc()
        else:
            a = np.asarray(x)
            if a.dtype not in _NP2RT:
                raise OpError(2, f"unsupported dtype {a.dtype}")
            if any(s < 0 for s in a.strides):
                a = np.ascontiguousarray(a)
            self.keep.append(a)
            d = _desc(a.ctypes.data, a.dtype, a.shape, [s // a.itemsize for s in a.strides], RTEN_DEVICE_HOST)
        self.keep.append(d)
        return C.byref(d)

    def out(self, into: Optional[DeviceTensor] = None):
        d = into.desc() if into is not None else RtenTensor()
        self.keep.append(d)
        return d

    def wrap(self, d: RtenTensor, into: Optional[DeviceTensor]) -> DeviceTensor:
        if into is not None:
            return into
        shape = tuple(d.shape[i] for i in range(d.ndim))
        strides = tuple(d.strides[i] for i in range(d.ndim))
        return DeviceTensor(self.ctx, d.data, shape, strides, _RT2NP[d.dtype], owner=True)


class Packed:
    """`PrepackedInput` (src/operator.rs:25-31)."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self.handle = ctx, handle

    def __del__(self):
        try:
            if self.handle and self.ctx.handle:
                self.ctx.lib.rten_b200_packed_free(self.ctx.handle, self.handle)
        except Exception:
            pass


def _ph(p: Optional[Packed]):
    return p.handle if p is not None else None


# =========================================================================================
# Operators
# =========================================================================================
class Gemm:
    """src/ops/matmul.rs:106-147; ONNX defaults alpha = beta = 1 (src/op_registry/onnx_registry.rs:1184-1196)."""

    def __init__(self, alpha=1.0, beta=1.0, transpose_a=False, transpose_b=False):
        self.alpha, self.beta, self.transpose_a, self.transpose_b = alpha, beta, transpose_a, transpose_b

    def run(self, ctx, a, b, c=None, out=None):
        A = _Args(ctx)
        o = A.out(out)
        ctx.check(ctx.lib.rten_b200_gemm(ctx.handle, A.t(a), A.t(b), A.t(c), self.alpha, self.beta, int(self.transpose_a),
                                         int(self.transpose_b), C.byref(o)))
        return A.wrap(o, out)


class MatMul:
    """src/ops/matmul.rs:387-434"""

    def prepack_inputs(self):
        return [1]

    def prepack(self, ctx, index, value):
        if index != 1:
            return None
        A = _Args(ctx)
        h = C.c_void_p()
        ctx.check(ctx.lib.rten_b200_prepack_b(ctx.handle, A.t(value), C.byref(h)))
        return Packed(ctx, h)

    def run(self, ctx, a, b, packed_b: Optional[Packed] = None, out=None):
        A = _Args(ctx)
        o = A.out(out)
        ctx.check(ctx.lib.rten_b200_matmul(ctx.handle, A.t(a), A.t(b), _ph(packed_b), None, 1.0, C.byref(o)))
        return A.wrap(o, out)


class FusedMatMul(MatMul):
    """src/ops/matmul.rs:459-507: optional `alpha`, row bias as third input.  `activation` / `residual`
    are this backend's epilogue extensions (0 = reference behaviour)."""

    def __init__(self, alpha: Optional[float] = None, activation: int = ACT_NONE):
        self.alpha, self.activation = alpha, activation

    def run(self, ctx, a, b, bias=None, packed_b: Optional[Packed] = None, residual=None, out=None):
        A = _Args(ctx)
        o = A.out(out)
        ctx.check(ctx.lib.rten_b200_matmul_ex(ctx.handle, A.t(a), A.t(b), _ph(packed_b), A.t(bias),
                                              1.0 if self.alpha is None else self.alpha, A.t(residual), self.activation,
                                              C.byref(o)))
        return A.wrap(o, out)


class MatMulInteger(MatMul):
    """src/ops/matmul.rs:649-697"""

    def run(self, ctx, a, b, a_zero_point=None, b_zero_point=None, packed_b: Optional[Packed] = None, out=None):
        A = _Args(ctx)
        o = A.out(out)
        ctx.check(ctx.lib.rten_b200_matmul_integer(ctx.handle, A.t(a), A.t(b), _ph(packed_b), A.t(a_zero_point),
                                                   A.t(b_zero_point), None, C.byref(o)))
        return A.wrap(o, out)


class MatMulIntegerToFloat(MatMul):
    """src/ops/matmul.rs:776-811 (inputs: a, b, a_zero_point, b_zero_point, scale)"""

    def __init__(self, activation: int = ACT_NONE):
        self.activation = activation

    def run(self, ctx, a, b, a_zero_point, b_zero_point, scale, packed_b: Optional[Packed] = None, out=None, bias=None,
            residual=None, scale_b=None, out_range=None):
        """`bias` / `residual` / `self.activation`: the Add / Add / Gelu nodes that follow the operator in a quantised
        transformer, folded into the epilogue with the same f32 roundings (rten_b200_matmul_integer_ex)."""
        if scale is None:
            raise OpError(4, "missing inputs")
        A = _Args(ctx)
        o = A.out(out)
        ctx.check(ctx.lib.rten_b200_matmul_integer_ex(ctx.handle, A.t(a), A.t(b), _ph(packed_b), A.t(a_zero_point),
                                                      A.t(b_zero_point), A.t(scale), A.t(scale_b), A.t(bias), A.t(residual),
                                                      self.activation, A.t(out_range), C.byref(o)))
        return A.wrap(o, out)


class QuantizedLinear:
    """[LayerNormalization] -> DynamicQuantizeLinear -> Mul -> MatMulIntegerToFloat -> Add(bias) -> Add(residual) -> activation
    as one call (rten_b200_quantized_linear): the skinny-M decode kernel for <= 16 rows, the operator chain otherwise;
    bit-identical to the separate operators either way."""

    def __init__(self, activation: int = ACT_NONE, ln_epsilon: Optional[float] = None):
        self.activation, self.ln_epsilon = activation, ln_epsilon

    def run(self, ctx, x, w, w_scale, packed_w: Optional[Packed] = None, w_zero_point=None, bias=None, residual=None,
            ln_scale=None, ln_bias=None, out=None):
        A = _Args(ctx)
        o = A.out(out)
        ctx.check(ctx.lib.rten_b200_quantized_linear(ctx.handle, A.t(x), A.t(ln_scale), A.t(ln_bias),
                                                     -1.0 if self.ln_epsilon is None else float(self.ln_epsilon), A.t(w), _ph(packed_w),
                                                     A.t(w_zero_point), A.t(w_scale), A.t(bias), A.t(residual), self.activation, C.byref(o)))
        return A.wrap(o, out)


class Attention:
    """src/ops/attention.rs:645-905 (ONNX `Attention`) on 4-D inputs; attributes as the reference's."""

    def __init__(self, is_causal=False, kv_num_heads=None, q_num_heads=None, scale: Optional[float] = None, softcap: float = 0.0):
        self.is_causal, self.kv_num_heads, self.q_num_heads, self.scale, self.softcap = is_causal, kv_num_heads, q_num_heads, scale, softcap

    def run(self, ctx, query, key, value, attn_mask=None, nonpad_kv_seqlen=None, new_key=None, new_value=None, out=None):
        """`new_key` / `new_value`: this step's key / value [batch, kv_heads, 1, head], appended to the caches `key` /
        `value` at position nonpad_kv_seqlen[b] - 1 by the same kernel (q_seq = 1 only)."""
        A = _Args(ctx)
        o = A.out(out)
        p = RtenAttentionParams(int(bool(self.is_causal)), int(self.q_num_heads or 0), int(self.kv_num_heads or 0),
                                float(self.scale) if self.scale else 0.0, float(self.softcap))
        ctx.check(ctx.lib.rten_b200_attention(ctx.handle, A.t(query), A.t(key), A.t(value), A.t(attn_mask), A.t(nonpad_kv_seqlen),
                                              C.byref(p), A.t(new_key), A.t(new_value), C.byref(o)))
        return A.wrap(o, out)


def _conv_params(padding, groups, strides, dilations) -> RtenConvParams:
    p = RtenConvParams()
    if isinstance(padding, str):
        if padding.lower() != "same":
            raise OpError(5, "unknown padding mode")
        p.auto_pad_same = 1
    else:
        pads = list(padding)
        if len(pads) == 2:  # 1-D [start, end]
            pads = [pads[0], pads[1], 0, 0]
        if len(pads) != 4:
            raise OpError(5, "Wrong number of pad values")
        for i in range(4):
            p.pads[i] = int(pads[i])
    p.groups = int(groups)
    p.n_strides = len(strides)
    p.n_dilations = len(dilations)
    for i in range(min(2, len(strides))):
        p.strides[i] = int(strides[i])
    for i in range(min(2, len(dilations))):
        p.dilations[i] = int(dilations[i])
    return p


class Conv:
    """src/ops/conv.rs:367-419: attributes groups, dilations, padding ('same' or [t,l,b,r]), strides."""

    def __init__(self, groups=1, dilations=(1, 1), padding=(0, 0, 0, 0), strides=(1, 1), activation: int = ACT_NONE):
        self.groups, self.dilations, self.padding, self.strides = groups, tuple(dilations), padding, tuple(strides)
        self.activation = activation

    def prepack(self, ctx, index, value):
        if index != 1:
            return None
        A = _Args(ctx)
        h = C.c_void_p()
        ctx.check(ctx.lib.rten_b200_prepack_conv_weight(ctx.handle, A.t(value), self.groups, C.byref(h)))
        return Packed(ctx, h)

    def run(self, ctx, x, w, bias=None, packed_w: Optional[Packed] = None, residual=None, out=None):
        A = _Args(ctx)
        o = A.out(out)
        p = _conv_params(self.padding, self.groups, self.strides, self.dilations)
        ctx.check(ctx.lib.rten_b200_conv2d_ex(ctx.handle, A.t(x), A.t(w), _ph(packed_w), A.t(bias), C.byref(p),
                                              A.t(residual), self.activation, C.byref(o)))
        return A.wrap(o, out)


class ConvInteger(Conv):
    """src/ops/conv.rs:477-533"""

    def run(self, ctx, x, w, x_zero_point=None, w_zero_point=None, packed_w: Optional[Packed] = None, out=None):
        A = _Args(ctx)
        o = A.out(out)
        p = _conv_params(self.padding, self.groups, self.strides, self.dilations)
        ctx.check(ctx.lib.rten_b200_conv_integer(ctx.handle, A.t(x), A.t(w), _ph(packed_w), A.t(x_zero_point),
                                                 A.t(w_zero_point), None, C.byref(p), C.byref(o)))
        return A.wrap(o, out)


class ConvIntegerToFloat(Conv):
    """src/ops/conv.rs:535-587"""

    def run(self, ctx, x, w, x_zero_point, w_zero_point, scale, packed_w: Optional[Packed] = None, out=None,
            bias=None, residual=None, scale_b=None, out_range=None):
        """`bias` / `residual` / `self.activation` = the Add(bias), Add(identity), Relu nodes that follow the operator in
        a quantised ResNet, executed in the epilogue with the same f32 roundings (rten_b200_conv_integer_ex)."""
        if scale is None:
            raise OpError(4, "missing inputs")
        A = _Args(ctx)
        o = A.out(out)
        p = _conv_params(self.padding, self.groups, self.strides, self.dilations)
        ctx.check(ctx.lib.rten_b200_conv_integer_ex(ctx.handle, A.t(x), A.t(w), _ph(packed_w), A.t(x_zero_point),
                                                    A.t(w_zero_point), A.t(scale), A.t(scale_b), C.byref(p), A.t(bias),
                                                    A.t(residual), self.activation, A.t(out_range), C.byref(o)))
        return A.wrap(o, out)


class Softmax:
    """src/ops/norm.rs:843-899; `in_place` = run_in_place on input 0."""

    def __init__(self, axis=-1, flush_nans_to_zero=False):
        self.axis, self.flush_nans_to_zero = axis, flush_nans_to_zero

    def run(self, ctx, x, in_place=False):
        A = _Args(ctx)
        into = x if (in_place and isinstance(x, DeviceTensor)) else None
        o = A.out(into)
        ctx.check(ctx.lib.rten_b200_softmax(ctx.handle, A.t(x), None, self.axis, int(self.flush_nans_to_zero), C.byref(o)))
        return A.wrap(o, into)


class AddSoftmax:
    """src/ops/attention.rs:72-165: the larger input is QK; the other is broadcast to it."""

    def __init__(self, flush_nans_to_zero=False):
        self.flush_nans_to_zero = flush_nans_to_zero

    def run(self, ctx, x, y, in_place=False):
        nx = x.size if isinstance(x, DeviceTensor) else np.asarray(x).size
        ny = y.size if isinstance(y, DeviceTensor) else np.asarray(y).size
        qk, m = (x, y) if nx > ny else (y, x)
        A = _Args(ctx)
        into = qk if (in_place and isinstance(qk, DeviceTensor)) else None
        o = A.out(into)
        ctx.check(ctx.lib.rten_b200_softmax(ctx.handle, A.t(qk), A.t(m), -1, int(self.flush_nans_to_zero), C.byref(o)))
        return A.wrap(o, into)


class LayerNormalization:
    """src/ops/norm.rs:531-569"""

    def __init__(self, axis=-1, epsilon: Optional[float] = None):
        self.axis, self.epsilon = axis, epsilon

    def run(self, ctx, x, scale, bias=None, out=None):
        A = _Args(ctx)
        o = A.out(out)
        ctx.check(ctx.lib.rten_b200_layer_norm(ctx.handle, A.t(x), A.t(scale), A.t(bias), self.axis,
                                               -1.0 if self.epsilon is None else float(self.epsilon), C.byref(o)))
        return A.wrap(o, out)


class _Unary:
    fn = ""

    def run(self, ctx, x, in_place=False):
        A = _Args(ctx)
        into = x if (in_place and isinstance(x, DeviceTensor)) else None
        o = A.out(into)
        ctx.check(self._call(ctx, A.t(x), C.byref(o)))
        return A.wrap(o, into)


class Erf(_Unary):
    """src/ops/unary_elementwise.rs:384-387"""

    def _call(self, ctx, x, o):
        return ctx.lib.rten_b200_erf(ctx.handle, x, o)


class Gelu(_Unary):
    """src/ops/unary_elementwise.rs:399-435"""

    def __init__(self, approximate=False):
        self.approximate = approximate

    def _call(self, ctx, x, o):
        return ctx.lib.rten_b200_gelu(ctx.handle, x, int(self.approximate), o)


class Relu(_Unary):
    def _call(self, ctx, x, o):
        return ctx.lib.rten_b200_relu(ctx.handle, x, o)


class DynamicQuantizeLinear:
    """src/ops/quantize.rs:436-468 -> (y u8, y_scale f32 scalar, y_zero_point u8 scalar)"""

    def run(self, ctx, x, comm: Optional["Comm"] = None, value_range=None, out=None):
        """`comm`: batch-sharded run -- the quantisation range is all-reduced over the ranks (min, max) first.
        `value_range`: i32[2] device tensor filled by the producer of x (`out_range=` of the *IntegerToFloat operators):
        the operator then skips its own min / max pass.  `out`: u8 destination view, e.g. the interior of a spatially
        pre-padded channels-last buffer (rows at arbitrary pitches)."""
        A = _Args(ctx)
        y, s, z = A.out(out), A.out(), A.out()
        ctx.check(ctx.lib.rten_b200_dynamic_quantize_linear_ranged(ctx.handle, A.t(x), A.t(value_range), C.byref(y), C.byref(s),
                                                                   C.byref(z), comm.handle if comm is not None else None))
        return A.wrap(y, out), A.wrap(s, None), A.wrap(z, None)

    @staticmethod
    def reset_ranges(ctx, ranges: "DeviceTensor"):
        """Re-arm a [n, 2] i32 tensor of producer-computed ranges (one launch) before the producers run."""
        d = ranges.desc()
        ctx.check(ctx.lib.rten_b200_range_reset(ctx.handle, C.byref(d)))


class Comm:
    """Cross-rank communicator of a batch-sharded run (rten_b200_comm_*; NCCL resolved at run time)."""

    def __init__(self, ctx: Context, unique_id: bytes, rank: int, world: int):
        assert len(unique_id) == 128
        self.ctx, self.rank, self.world = ctx, rank, world
        h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128)
        ctx.check(ctx.lib.rten_b200_comm_create(ctx.handle, buf, rank, world, C.byref(h)))
        self.handle = h

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        st = _lib.load().rten_b200_comm_unique_id(buf)
        if st != 0:
            raise OpError(st, "libnccl.so.2 could not be loaded")
        return buf.raw

    @property
    def uses_peer_memory(self) -> bool:
        """True: quantisation ranges travel through NVLink peer mailboxes (one kernel per exchange); False: NCCL."""
        return bool(self.ctx.lib.rten_b200_comm_uses_peer_memory(self.handle))

    def timeouts(self) -> int:
        return int(self.ctx.lib.rten_b200_comm_timeouts(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.rten_b200_comm_destroy(self.handle)
            self.handle = None


class Add:
    def run(self, ctx, a, b, out=None):
        A = _Args(ctx)
        o = A.out(out)
        ctx.check(ctx.lib.rten_b200_add(ctx.handle, A.t(a), A.t(b), C.byref(o)))
        return A.wrap(o, out)


class Mul:
    """src/ops/binary_elementwise.rs Mul (f32, broadcasting)."""

    def run(self, ctx, a, b, out=None):
        A = _Args(ctx)
        o = A.out(out)
        ctx.check(ctx.lib.rten_b200_mul(ctx.handle, A.t(a), A.t(b), C.byref(o)))
        return A.wrap(o, out)


class MaxPool:
    """src/ops/pooling.rs MaxPool: kernel_size, padding [t,l,b,r], strides."""

    def __init__(self, kernel_size, padding=(0, 0, 0, 0), strides=(1, 1)):
        self.kernel_size, self.padding, self.strides = tuple(kernel_size), tuple(padding), tuple(strides)

    def run(self, ctx, x, out=None):
        A = _Args(ctx)
        o = A.out(out)
        k = (C.c_int32 * 2)(*self.kernel_size)
        p = (C.c_int32 * 4)(*self.padding)
        s = (C.c_int32 * 2)(*self.strides)
        ctx.check(ctx.lib.rten_b200_max_pool(ctx.handle, A.t(x), k, p, s, C.byref(o)))
        return A.wrap(o, out)


class GlobalAveragePool:
    def run(self, ctx, x, out=None):
        A = _Args(ctx)
        o = A.out(out)
        ctx.check(ctx.lib.rten_b200_global_average_pool(ctx.handle, A.t(x), C.byref(o)))
        return A.wrap(o, out)


class ScatterRows:
    """table[indices[r], :] = updates[r, :] in place (2-D f32 table, distinct i32 indices)."""

    def run(self, ctx, table: "DeviceTensor", indices, updates):
        A = _Args(ctx)
        t = table.desc()
        ctx.check(ctx.lib.rten_b200_scatter_rows(ctx.handle, C.byref(t), A.t(indices), A.t(updates)))
        return table


class GatherRows:
    """Gather(axis=0) of a 2-D table with i32 indices (embedding lookup; src/ops/gather.rs)."""

    def run(self, ctx, table, indices, out=None):
        A = _Args(ctx)
        o = A.out(out)
        ctx.check(ctx.lib.rten_b200_gather_rows(ctx.handle, A.t(table), A.t(indices), C.byref(o)))
        return A.wrap(o, out)
