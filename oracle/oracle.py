"""Operator-level CPU oracle for the rten hot path (numpy host logic over librten_oracle.so).

TEST INFRASTRUCTURE ONLY (see rten_oracle.c header).  Importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never from
rten_b200/.

Each function restates one reference operator and cites the reference file:line
(robertknight/rten @ c7f7bad).  Error behaviour mirrors `OpError`
(src/operator.rs:116-144) with the reference's message strings.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librten_oracle.so")


def build(force: bool = False) -> str:
    """Compile librten_oracle.so with oracle/Makefile (gcc only; no reference build system)."""
    src = os.path.join(_HERE, "rten_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "librten_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        f32p, u8p, i8p, i32p, f64p = (C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int8),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_double))
        sz, pd, ip = C.c_size_t, C.c_ssize_t, C.POINTER(C.c_int)
        u64p = C.POINTER(C.c_uint64)
        sig = {
            "rto_rng_u64": (None, [u64p, u64p, sz]),
            "rto_rng_f32": (None, [u64p, f32p, sz]),
            "rto_rng_u8": (None, [u64p, u8p, sz, C.c_int]),
            "rto_rng_i8": (None, [u64p, i8p, sz, C.c_int]),
            "rto_rng_i32": (None, [u64p, i32p, sz]),
            "rto_exp": (None, [f32p, f32p, sz]),
            "rto_erf": (None, [f32p, f32p, sz]),
            "rto_gelu": (None, [f32p, f32p, sz]),
            "rto_approx_gelu": (None, [f32p, f32p, sz]),
            "rto_tanh": (None, [f32p, f32p, sz]),
            "rto_relu": (None, [f32p, f32p, sz]),
            "rto_sum": (C.c_float, [f32p, sz]),
            "rto_softmax": (None, [f32p, f32p, f32p, sz, sz, C.c_int]),
            "rto_layer_norm": (None, [f32p, f32p, sz, sz, f32p, C.c_float, f32p, C.c_float, C.c_float]),
            "rto_gemm_f32": (None, [sz, sz, sz, f32p, pd, pd, f32p, pd, pd, f32p, C.c_float, C.c_float, f32p, C.c_int]),
            "rto_gemm_f64": (None, [sz, sz, sz, f32p, pd, pd, f32p, pd, pd, f64p, f64p]),
            "rto_gemm_u8i8": (None, [sz, sz, sz, u8p, pd, pd, i8p, pd, pd, i32p, u8p, i8p]),
            "rto_cast_scale": (None, [i32p, f32p, sz, sz, f32p, sz]),
            "rto_conv_f32": (None, [f32p, f32p, f32p, f32p] + [sz] * 9 + [ip, ip, ip, sz]),
            "rto_conv_u8i8": (None, [i8p, u8p, i32p] + [sz] * 9 + [ip, ip, ip, sz, C.c_int8, u8p]),
            "rto_quantize_u8": (None, [f32p, u8p, sz, C.c_float, C.c_uint8]),
            "rto_dynamic_quantize_linear": (None, [f32p, sz, u8p, f32p, u8p]),
            "rto_maxpool2d": (None, [f32p, f32p] + [sz] * 8 + [ip, ip]),
            "rto_global_avgpool": (None, [f32p, f32p, sz, sz]),
            "rto_add": (None, [f32p, f32p, f32p, sz]),
            "rto_num_threads": (C.c_int, []),
            "rto_set_num_threads": (None, [C.c_int]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray], ct):
    if a is None:
        return C.cast(None, C.POINTER(ct))
    return a.ctypes.data_as(C.POINTER(ct))


def _ints(v: Sequence[int]):
    return (C.c_int * len(v))(*[int(x) for x in v])


class OpError(Exception):
    """Mirror of `OpError` (src/operator.rs:116-144): kind + the reference's static message."""

    def __init__(self, kind: str, msg: str = ""):
        super().__init__(f"{kind}: {msg}" if msg else kind)
        self.kind = kind
        self.msg = msg


# --------------------------------------------------------------------------------------
# RNG + comparison rule
# --------------------------------------------------------------------------------------
class XorShiftRng:
    """rten-tensor/src/rng.rs:6-66; reduced range: rten-gemm/src/reduced_range_rng.rs:37-57."""

    def __init__(self, seed: int):
        self._state = C.c_uint64(seed)

    def _run(self, fn, dtype, n, *extra):
        out = np.empty(int(n), dtype=dtype)
        ct = {np.float32: C.c_float, np.uint8: C.c_uint8, np.int8: C.c_int8, np.int32: C.c_int32,
              np.uint64: C.c_uint64}[dtype]
        fn(C.byref(self._state), _p(out, ct), out.size, *extra)
        return out

    def u64(self, n):
        return self._run(lib().rto_rng_u64, np.uint64, n)

    def f32(self, shape):
        return self._run(lib().rto_rng_f32, np.float32, int(np.prod(shape))).reshape(shape)

    def u8(self, shape, reduce_range=False):
        return self._run(lib().rto_rng_u8, np.uint8, int(np.prod(shape)), int(reduce_range)).reshape(shape)

    def i8(self, shape, reduce_range=False):
        return self._run(lib().rto_rng_i8, np.int8, int(np.prod(shape)), int(reduce_range)).reshape(shape)

    def i32(self, shape):
        return self._run(lib().rto_rng_i32, np.int32, int(np.prod(shape))).reshape(shape)

    def uniform(self, shape, lo=-1.0, hi=1.0):
        """U(lo,hi) from next_f32 (SURVEY.md 8d synthetic-input recipe)."""
        return (self.f32(shape) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)


def expect_equal(a, b, atol=1e-8, rtol=1e-5) -> bool:
    """rten-tensor/src/test_util.rs:47-92: f32 `a==b or |a-b| <= atol + rtol*|b|`; ints exact."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        return False
    if np.issubdtype(a.dtype, np.integer):
        return bool(np.array_equal(a, b))
    with np.errstate(invalid="ignore"):
        ok = (a == b) | (np.abs(a - b) <= atol + rtol * np.abs(b))
    return bool(np.all(ok))


# --------------------------------------------------------------------------------------
# Elementwise
# --------------------------------------------------------------------------------------
def _unary(fn, x, out=None):
    """`out` (may be `x` itself): run into an existing buffer, as the reference's in-place operators do."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = out if out is not None else np.empty_like(x)
    assert y.shape == x.shape and y.dtype == np.float32 and y.flags.c_contiguous
    fn(_p(x, C.c_float), _p(y, C.c_float), x.size)
    return y


def exp(x):
    """rten-vecmath/src/exp.rs:61-127"""
    return _unary(lib().rto_exp, x)


def erf(x):
    """src/ops/unary_elementwise.rs:384-387 -> rten-vecmath/src/erf.rs:23-56"""
    return _unary(lib().rto_erf, x)


def gelu(x, approximate: bool = False):
    """src/ops/unary_elementwise.rs:399-435 -> erf.rs:65-76 (erf) / :86-100 (tanh approx)"""
    return _unary(lib().rto_approx_gelu if approximate else lib().rto_gelu, x)


def tanh(x):
    """rten-vecmath/src/tanh.rs:12-66"""
    return _unary(lib().rto_tanh, x)


def relu(x, out=None):
    return _unary(lib().rto_relu, x, out)


def add(a, b, out=None):
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    shape = np.broadcast_shapes(a.shape, b.shape)
    a = np.ascontiguousarray(np.broadcast_to(a, shape))
    b = np.ascontiguousarray(np.broadcast_to(b, shape))
    y = out if out is not None else np.empty(shape, dtype=np.float32)
    assert y.shape == tuple(shape) and y.dtype == np.float32 and y.flags.c_contiguous
    lib().rto_add(_p(a, C.c_float), _p(b, C.c_float), _p(y, C.c_float), y.size)
    return y


class Arena:
    """Output buffers recycled from one model pass to the next, the role `BufferPool` (src/buffer_pool.rs) plays in
    the reference: without it every operator output is a fresh multi-megabyte allocation whose first-touch page
    faults dominate a many-core CPU run."""

    def __init__(self):
        self.bufs = {}

    def get(self, key, shape, dtype=np.float32):
        a = self.bufs.get(key)
        if a is None or a.shape != tuple(shape) or a.dtype != dtype:
            a = np.empty(shape, dtype)
            self.bufs[key] = a
        return a


# --------------------------------------------------------------------------------------
# Softmax / AddSoftmax / LayerNormalization
# --------------------------------------------------------------------------------------
def _resolve_axis(ndim: int, axis: int) -> int:
    if axis < -ndim or axis >= ndim:
        raise OpError("InvalidValue", "Axis is invalid")
    return axis % ndim if ndim else 0


def softmax(x, axis: int = -1, flush_nans_to_zero: bool = False, mask=None):
    """src/ops/norm.rs:705-755,825-899 (normalize_lanes: move axis last, contiguous lanes) ->
    rten-vecmath/src/softmax.rs:60-101."""
    x = np.asarray(x, dtype=np.float32)
    if x.ndim == 0:
        raise OpError("InvalidValue", "Axis is invalid")
    ax = _resolve_axis(x.ndim, axis)
    xm = np.ascontiguousarray(np.moveaxis(x, ax, -1))
    m = None
    if mask is not None:
        m = np.ascontiguousarray(np.moveaxis(np.broadcast_to(np.asarray(mask, np.float32), x.shape), ax, -1))
    n = xm.shape[-1]
    rows = xm.size // n if n else 0
    y = np.empty_like(xm)
    if xm.size:
        lib().rto_softmax(_p(xm, C.c_float), _p(m, C.c_float), _p(y, C.c_float), rows, n, int(flush_nans_to_zero))
    return np.ascontiguousarray(np.moveaxis(y, -1, ax))


def add_softmax(x, y, flush_nans_to_zero: bool = False):
    """src/ops/attention.rs:30-121: the larger input is QK, the other is broadcast to it, lane-wise
    `qk += m` then Softmax(axis=-1)."""
    x = np.asarray(x, np.float32)
    y = np.asarray(y, np.float32)
    qk, m = (x, y) if x.size > y.size else (y, x)
    try:
        shape = np.broadcast_shapes(qk.shape, m.shape)
    except ValueError:
        raise OpError("IncompatibleInputShapes", "Cannot broadcast inputs")
    qk = np.broadcast_to(qk, shape)
    return softmax(qk, -1, flush_nans_to_zero, mask=m)


def layer_norm(x, scale, bias=None, axis: int = -1, epsilon: Optional[float] = None):
    """src/ops/norm.rs:456-529 (+ normalize_slice :103-161)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    eps = 1e-5 if epsilon is None else float(epsilon)
    ax = _resolve_axis(x.ndim, axis)
    nshape = x.shape[ax:]
    scale = np.asarray(scale, np.float32)
    g = None
    gs = 1.0
    if scale.size == 1:
        gs = float(scale.reshape(-1)[0])
    else:
        try:
            g = np.ascontiguousarray(np.broadcast_to(scale, nshape))
        except ValueError:
            raise OpError("InvalidValue", "`scale` is not broadcastable to normalized axes of input")
    b = None
    bs = 0.0
    if bias is not None:
        bias = np.asarray(bias, np.float32)
        if bias.size == 1:
            bs = float(bias.reshape(-1)[0])
        else:
            try:
                b = np.ascontiguousarray(np.broadcast_to(bias, nshape))
            except ValueError:
                raise OpError("InvalidValue", "`bias` is not broadcastable to normalized axes of input")
    n = int(np.prod(nshape))
    rows = x.size // n if n else 0
    y = np.empty_like(x)
    if x.size:
        lib().rto_layer_norm(_p(x, C.c_float), _p(y, C.c_float), rows, n, _p(g, C.c_float), gs, _p(b, C.c_float), bs, eps)
    return y


# --------------------------------------------------------------------------------------
# GEMM level (rten-gemm GemmExecutor::gemm / gemm_uninit)
# --------------------------------------------------------------------------------------
def _strides_el(a: np.ndarray):
    return [s // a.itemsize for s in a.strides]


def gemm_f32(a, b, c=None, alpha=1.0, beta=0.0, bias=None, bias_kind: Optional[str] = None):
    """rten-gemm/src/lib.rs:263-372,794-1093.  a: [M,K], b: [K,N] any strides; c: [M,N] initial
    output (read only when beta != 0).  bias_kind in {None,'row','column'}."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    M, K = a.shape
    K2, N = b.shape
    if K != K2:
        raise OpError("GemmError", "KSizeMismatch")
    if bias is not None:
        bias = np.ascontiguousarray(bias, np.float32)
        if bias.size != (N if bias_kind == "row" else M):
            raise OpError("GemmError", "WrongBiasSize")
    if c is None or beta == 0.0:
        out = np.full((M, N), np.nan, dtype=np.float32)  # poison: beta==0 must not read C
    else:
        out = np.ascontiguousarray(c, np.float32).copy()
        if out.shape != (M, N):
            raise OpError("GemmError", "OutputSizeMismatch")
    ars, acs = _strides_el(a) if a.size else (K, 1)
    brs, bcs = _strides_el(b) if b.size else (N, 1)
    kind = {None: 0, "row": 1, "column": 2}[bias_kind if bias is not None else None]
    if M and N:
        lib().rto_gemm_f32(M, N, K, _p(a, C.c_float), ars, acs, _p(b, C.c_float), brs, bcs, _p(out, C.c_float),
                           float(alpha), float(beta), _p(bias, C.c_float), kind)
    return out


def gemm_f64(a, b):
    """float64 truth + sum|a||b| (tolerance bound for the TF32 GPU path)."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    M, K = a.shape
    _, N = b.shape
    c = np.zeros((M, N), np.float64)
    ca = np.zeros((M, N), np.float64)
    ars, acs = _strides_el(a) if a.size else (K, 1)
    brs, bcs = _strides_el(b) if b.size else (N, 1)
    if M and N:
        lib().rto_gemm_f64(M, N, K, _p(a, C.c_float), ars, acs, _p(b, C.c_float), brs, bcs, _p(c, C.c_double), _p(ca, C.c_double))
    return c, ca


def gemm_u8i8(a, b, a_zp=None, b_zp=None):
    """rten-gemm u8 x i8 -> i32 with per-row / per-column zero points (simd_generic.rs:576-780)."""
    a = np.asarray(a, np.uint8)
    b = np.asarray(b, np.int8)
    M, K = a.shape
    K2, N = b.shape
    if K != K2:
        raise OpError("GemmError", "KSizeMismatch")
    if a_zp is not None:
        a_zp = np.ascontiguousarray(a_zp, np.uint8)
        if a_zp.size != M:
            raise OpError("GemmError", "WrongQuantParamSize")
    if b_zp is not None:
        b_zp = np.ascontiguousarray(b_zp, np.int8)
        if b_zp.size != N:
            raise OpError("GemmError", "WrongQuantParamSize")
    out = np.zeros((M, N), np.int32)
    ars, acs = _strides_el(a) if a.size else (K, 1)
    brs, bcs = _strides_el(b) if b.size else (N, 1)
    if M and N:
        lib().rto_gemm_u8i8(M, N, K, _p(a, C.c_uint8), ars, acs, _p(b, C.c_int8), brs, bcs, _p(out, C.c_int32),
                            _p(a_zp, C.c_uint8), _p(b_zp, C.c_int8))
    return out


# --------------------------------------------------------------------------------------
# Gemm / MatMul / FusedMatMul operators
# --------------------------------------------------------------------------------------
def gemm_op(a, b, c=None, alpha=1.0, beta=1.0, trans_a=False, trans_b=False):
    """ONNX Gemm: src/ops/matmul.rs:32-104 (C broadcast into the output, then GEMM with beta)."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    if a.ndim != 2 or b.ndim != 2:
        raise OpError("InvalidValue", "input must have 2 dims")
    if trans_a:
        a = a.T
    if trans_b:
        b = b.T
    if a.shape[1] != b.shape[0]:
        raise OpError("IncompatibleInputShapes", "Columns of first matrix does not match rows of second matrix")
    out_shape = (a.shape[0], b.shape[1])
    if c is not None and beta != 0.0:
        c = np.asarray(c, np.float32)
        try:
            cb = np.broadcast_to(c, out_shape)
        except ValueError:
            raise OpError("IncompatibleInputShapes", "Cannot broadcast c to output shape")
        return gemm_f32(a, b, c=cb, alpha=alpha, beta=beta)
    return gemm_f32(a, b, alpha=alpha, beta=0.0)


def _matmul_core(a, b, gemm2d, out_dtype, row_quant=None):
    """src/ops/matmul.rs:208-385: numpy.matmul broadcasting; [A.., M, K] x [K, N] is flattened to one
    [A*M, K] GEMM (:266-297) with row zero points cycled (:272-280)."""
    if a.ndim < 1 or b.ndim < 1:
        raise OpError("InvalidValue", "Inputs must have >= 1 dimensions")
    a_is_vec = a.ndim == 1
    b_is_vec = b.ndim == 1
    if a_is_vec:
        a = a[None, :]
    if b_is_vec:
        b = b[:, None]
    M, K = a.shape[-2:]
    K2, N = b.shape[-2:]
    if K != K2:
        raise OpError("IncompatibleInputShapes", "Columns of first matrix does not match rows of second matrix")
    try:
        prefix = np.broadcast_shapes(a.shape[:-2], b.shape[:-2])
    except ValueError:
        raise OpError("IncompatibleInputShapes", "Cannot broadcast shapes")
    out_shape = tuple(prefix) + (M, N)
    na = int(np.prod(a.shape[:-2]))
    nb = int(np.prod(b.shape[:-2]))
    if na > 1 and nb == 1:
        a2 = np.ascontiguousarray(a).reshape(na * M, K)
        rq = None if row_quant is None else np.resize(row_quant, na * M)
        out = gemm2d(a2, b.reshape(K, N) if b.ndim > 2 else b, rq).reshape(out_shape)
    elif int(np.prod(out_shape)) == 0:
        out = np.zeros(out_shape, out_dtype)
    else:
        ab = np.broadcast_to(a, tuple(prefix) + (M, K)).reshape((-1, M, K))
        bb = np.broadcast_to(b, tuple(prefix) + (K, N)).reshape((-1, K, N))
        out = np.stack([gemm2d(ab[i], bb[i], row_quant) for i in range(ab.shape[0])]).reshape(out_shape)
    if a_is_vec:
        out = np.squeeze(out, axis=-2)
    if b_is_vec:
        out = np.squeeze(out, axis=-1)
    return out


def matmul(a, b, bias=None, alpha: Optional[float] = None):
    """MatMul (src/ops/matmul.rs:390-408) / FusedMatMul (:462-507): optional row bias over N and
    alpha applied to the product before the bias (G13)."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    if bias is not None:
        bias = np.asarray(bias, np.float32)
        if bias.ndim != 1:
            raise OpError("InputCastFailed", "bias must be a vector")
    al = 1.0 if alpha is None else float(alpha)

    def g(a2, b2, _rq):
        return gemm_f32(a2, b2, alpha=al, beta=0.0, bias=bias, bias_kind="row" if bias is not None else None)

    return _matmul_core(a, b, g, np.float32)


def shift_cast_to_u8(x):
    """src/shift_cast.rs:39-50: i8 -> u8 by XOR 0x80 (no-op for u8)."""
    x = np.asarray(x)
    if x.dtype == np.uint8:
        return x
    return (x.view(np.uint8) ^ np.uint8(0x80)).astype(np.uint8)


def shift_cast_to_i8(x):
    x = np.asarray(x)
    if x.dtype == np.int8:
        return x
    return (x ^ np.uint8(0x80)).view(np.int8)


def _zero_point_to_vec(zp, expected_len, dtype):
    """src/ops/matmul.rs:513-533"""
    if zp is None:
        return np.zeros(expected_len, dtype)
    zp = np.asarray(zp, dtype)
    if zp.ndim == 0:
        return np.full(expected_len, zp, dtype)
    if zp.ndim == 1:
        if zp.shape[0] != expected_len:
            raise OpError("InvalidValue", "Zero point has incorrect size")
        return zp
    raise OpError("UnsupportedValue", "Only scalar or vector zero points are supported")


def matmul_integer(a, b, a_zero_point=None, b_zero_point=None):
    """src/ops/matmul.rs:582-647: all four u8/i8 combinations are normalised to u8 x i8 by XOR 0x80 on
    data and zero points (VNNI host: `may_saturate()` is false, so the plain sign flip path)."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.dtype not in (np.uint8, np.int8) or b.dtype not in (np.uint8, np.int8):
        raise OpError("UnsupportedType")
    a_rows = a.shape[-2] if a.ndim > 1 else 1
    b_cols = b.shape[-1] if b.ndim > 1 else 1
    az = _zero_point_to_vec(a_zero_point, a_rows, a.dtype)
    bz = _zero_point_to_vec(b_zero_point, b_cols, b.dtype)
    a8, az8 = shift_cast_to_u8(a), shift_cast_to_u8(az)
    b8, bz8 = shift_cast_to_i8(b), shift_cast_to_i8(bz)

    def g(a2, b2, rq):
        return gemm_u8i8(a2, b2, a_zp=rq, b_zp=bz8)

    return _matmul_core(a8, b8, g, np.int32, row_quant=az8)


def cast_scale(data, scale):
    """src/ops/matmul.rs:704-773"""
    data = np.ascontiguousarray(data, np.int32)
    scale = np.asarray(scale, np.float32)
    if scale.ndim > 1:
        raise OpError("InvalidValue", "scale should have rank 0 or 1")
    s = np.ascontiguousarray(scale.reshape(-1))
    if s.size != 1 and data.shape[-1] != s.size:
        raise OpError("IncompatibleInputShapes", "Scale length does not match tensor columns")
    out = np.empty(data.shape, np.float32)
    cols = data.shape[-1] if data.ndim else 1
    rows = data.size // cols if cols else 0
    if data.size:
        lib().rto_cast_scale(_p(data, C.c_int32), _p(out, C.c_float), rows, cols, _p(s, C.c_float), s.size)
    return out


def matmul_integer_to_float(a, b, a_zero_point, b_zero_point, scale):
    """src/ops/matmul.rs:776-811"""
    return cast_scale(matmul_integer(a, b, a_zero_point, b_zero_point), scale)


# --------------------------------------------------------------------------------------
# Conv / ConvInteger
# --------------------------------------------------------------------------------------
def _axis_out_and_pad(in_size, k, stride, pad, dilation):
    """src/ops/pooling.rs:63-123 (Floor rounding)"""
    if dilation <= 0:
        raise OpError("InvalidValue", "Dilations must be > 0")
    if k <= 0:
        raise OpError("InvalidValue", "Kernel size must be > 0")
    if stride <= 0:
        raise OpError("InvalidValue", "Strides must be > 0")
    if pad == "same":
        out = -(-in_size // stride)
        total = max((out - 1) * stride + (k - 1) * dilation + 1 - in_size, 0)
        return out, total // 2, -(-total // 2)
    ps, pe = pad
    padded = in_size + ps + pe
    dk = k + (k - 1) * (dilation - 1)
    if padded < dk:
        raise OpError("InvalidValue", "Input too small for kernel size")
    return (padded - dilation * (k - 1) - 1) // stride + 1, ps, pe


def conv_output_size(in_hw, k_hw, strides, padding, dilations):
    """padding: 'same' or [top, left, bottom, right] (src/ops/pooling.rs:139-159)."""
    if padding == "same":
        ph = pw = "same"
    else:
        if len(padding) != 4:
            raise OpError("InvalidValue", "Wrong number of pad values")
        ph, pw = (padding[0], padding[2]), (padding[1], padding[3])
    oh, pt, pb = _axis_out_and_pad(in_hw[0], k_hw[0], strides[0], ph, dilations[0])
    ow, pl, pr = _axis_out_and_pad(in_hw[1], k_hw[1], strides[1], pw, dilations[1])
    return oh, ow, [pt, pl, pb, pr]


def _conv_checks(x, w, groups, strides, dilations):
    if x.ndim != 4:
        raise OpError("InvalidValue", "input must have 4 dims (NCHW)")
    if w.ndim != 4:
        raise OpError("InvalidValue", "input must have 4 dims (OCHW)")
    if len(strides) != 2:
        raise OpError("InvalidValue", "expected 2 stride values")
    if len(dilations) != 2:
        raise OpError("InvalidValue", "expected 2 dilation values")
    B, Cin, H, W = x.shape
    O, kc, kh, kw = w.shape
    if groups == 0:
        raise OpError("InvalidValue", "Group count must be > 0")
    if Cin % groups != 0:
        raise OpError("InvalidValue", "Input channel count not divisible by groups")
    if Cin // groups != kc:
        raise OpError("IncompatibleInputShapes", "Input channels (per group) does not match kernel input channels")
    if O % groups != 0:
        raise OpError("InvalidValue", "Output channel count not divisible by groups")


def conv(x, w, bias=None, padding=(0, 0, 0, 0), groups=1, strides=(1, 1), dilations=(1, 1), out=None):
    """src/ops/conv.rs:124-365 (f32, NCHW x OIHW -> NCHW, column bias over out channels)."""
    x = np.ascontiguousarray(x, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    if x.ndim == 3:  # 1-D conv via 2-D (conv.rs:142-185)
        if w.ndim != 3:
            raise OpError("InvalidValue", "input must have 3 dims (OCW)")
        if len(strides) != 1:
            raise OpError("InvalidValue", "expected 1 stride value")
        if len(dilations) != 1:
            raise OpError("InvalidValue", "expected 1 dilation value")
        pad2 = padding if padding == "same" else [0, padding[0], 0, padding[1]]
        y = conv(x[:, :, None, :], w[:, :, None, :], bias, pad2, groups, (1, strides[0]), (1, dilations[0]))
        return y.reshape(y.shape[0], y.shape[1], y.shape[3])
    _conv_checks(x, w, groups, strides, dilations)
    B, Cin, H, W = x.shape
    O, _, kh, kw = w.shape
    if bias is not None:
        bias = np.ascontiguousarray(bias, np.float32)
        if bias.shape != (O,):
            raise OpError("IncompatibleInputShapes", "bias.size(0) != out_channels")
    oh, ow, pads = conv_output_size((H, W), (kh, kw), strides, padding, dilations)
    y = out if out is not None else np.empty((B, O, oh, ow), np.float32)
    assert y.shape == (B, O, oh, ow) and y.dtype == np.float32 and y.flags.c_contiguous
    if y.size:
        lib().rto_conv_f32(_p(x, C.c_float), _p(w, C.c_float), _p(bias, C.c_float), _p(y, C.c_float),
                           B, Cin, H, W, O, kh, kw, oh, ow, _ints(pads), _ints(strides), _ints(dilations), groups)
    return y


def conv_integer(x, w, x_zero_point=None, w_zero_point=None, padding=(0, 0, 0, 0), groups=1,
                 strides=(1, 1), dilations=(1, 1)):
    """src/ops/conv.rs:421-475.  x: u8|i8 NCHW, w: u8|i8 OIHW; x_zp scalar, w_zp scalar or [O].
    Padded taps follow the production path (literal 0 in the shifted-i8 domain, G3)."""
    x = np.ascontiguousarray(x)
    w = np.ascontiguousarray(w)
    if x.dtype not in (np.uint8, np.int8) or w.dtype not in (np.uint8, np.int8):
        raise OpError("UnsupportedType")
    O = w.shape[0] if w.ndim >= 1 else 0
    if x_zero_point is None:
        xz = np.zeros((), x.dtype)
    else:
        xz = np.asarray(x_zero_point, x.dtype)
        if xz.size != 1:
            raise OpError("InvalidValue", "input zero point must be a scalar")
        xz = xz.reshape(())
    wz = _zero_point_to_vec(w_zero_point, O, w.dtype)
    xs = np.ascontiguousarray(shift_cast_to_i8(x))
    xzs = shift_cast_to_i8(xz.reshape(1))[0]
    ws = np.ascontiguousarray(shift_cast_to_u8(w))
    wzs = np.ascontiguousarray(shift_cast_to_u8(wz))
    _conv_checks(xs, ws, groups, strides, dilations)
    B, Cin, H, W = xs.shape
    _, _, kh, kw = ws.shape
    oh, ow, pads = conv_output_size((H, W), (kh, kw), strides, padding, dilations)
    y = np.empty((B, O, oh, ow), np.int32)
    if y.size:
        lib().rto_conv_u8i8(_p(xs, C.c_int8), _p(ws, C.c_uint8), _p(y, C.c_int32), B, Cin, H, W, O, kh, kw, oh, ow,
                            _ints(pads), _ints(strides), _ints(dilations), groups, int(xzs), _p(wzs, C.c_uint8))
    return y


def conv_integer_to_float(x, w, x_zero_point, w_zero_point, scale, **kw):
    """src/ops/conv.rs:535-587: scale must be a scalar."""
    scale = np.asarray(scale, np.float32)
    if scale.size != 1:
        raise OpError("InvalidValue", "scale should be a scalar")
    return cast_scale(conv_integer(x, w, x_zero_point, w_zero_point, **kw), scale.reshape(()))


# --------------------------------------------------------------------------------------
# Quantisation
# --------------------------------------------------------------------------------------
def dynamic_quantize_linear(x) -> Tuple[np.ndarray, np.float32, np.uint8]:
    """src/ops/quantize.rs:352-434 -> (y u8, scale f32 scalar, zero_point u8 scalar)."""
    x = np.ascontiguousarray(x, np.float32)
    y = np.zeros(x.shape, np.uint8)
    scale = C.c_float(1.0)
    zp = C.c_uint8(0)
    lib().rto_dynamic_quantize_linear(_p(x, C.c_float), x.size, _p(y, C.c_uint8), C.byref(scale), C.byref(zp))
    return y, np.float32(scale.value), np.uint8(zp.value)


def quantize_u8(x, inv_scale, zero_point):
    """rten-vecmath/src/quantize.rs:38-77"""
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty(x.shape, np.uint8)
    lib().rto_quantize_u8(_p(x, C.c_float), _p(y, C.c_uint8), x.size, float(inv_scale), int(zero_point))
    return y


# --------------------------------------------------------------------------------------
# Residency glue
# --------------------------------------------------------------------------------------
def max_pool(x, kernel, padding=(0, 0, 0, 0), strides=(1, 1)):
    x = np.ascontiguousarray(x, np.float32)
    B, Cc, H, W = x.shape
    oh, ow, pads = conv_output_size((H, W), kernel, strides, padding, (1, 1))
    y = np.empty((B, Cc, oh, ow), np.float32)
    lib().rto_maxpool2d(_p(x, C.c_float), _p(y, C.c_float), B, Cc, H, W, kernel[0], kernel[1], oh, ow, _ints(pads), _ints(strides))
    return y


def global_average_pool(x):
    x = np.ascontiguousarray(x, np.float32)
    B, Cc, H, W = x.shape
    y = np.empty((B, Cc, 1, 1), np.float32)
    lib().rto_global_avgpool(_p(x, C.c_float), _p(y, C.c_float), B * Cc, H * W)
    return y


def num_threads() -> int:
    return int(lib().rto_num_threads())


def usable_cores() -> int:
    """Cores this process may actually use: its affinity mask, capped by the cgroup CPU quota (a container can see 128
    CPUs and be throttled to 16 -- running 128 spinning threads there is 10x slower than running 16)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    quota = None
    try:
        q, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = float(q) / float(period)
    except Exception:
        try:
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            period = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / period
        except Exception:
            pass
    if quota is not None:
        n = max(1, min(n, int(quota + 0.5)))
    return n


def use_all_cores() -> int:
    """Use every core this process may run on (torchrun sets OMP_NUM_THREADS=1 for its workers)."""
    lib().rto_set_num_threads(usable_cores())
    return num_threads()
