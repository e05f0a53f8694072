"""Pin the CPU oracle against every golden vector / known-answer test the reference's own tests hold
for the hot path.  Values are the literals asserted in the cited reference tests."""
import numpy as np
import pytest


def eq_1e4(a, b):
    """src/ops/mod.rs:407-412 expect_eq_1e4: atol 1e-4, rtol 0."""
    return np.all(np.abs(np.asarray(a) - np.asarray(b)) <= 1e-4)


# ---- RNG: rten-tensor/src/rng.rs:73-126 --------------------------------------------------
def test_rng_goldens(oracle):
    r = oracle.XorShiftRng(1234)
    np.testing.assert_array_equal(
        r.f32(10),
        np.array([7.2381226e-8, 0.12971127, 0.44675463, 6.69676e-5, 0.44387037, 0.24518594, 0.84056354,
                  0.9960614, 0.32433507, 0.9239961], np.float32))
    np.testing.assert_array_equal(oracle.XorShiftRng(1234).i8(10), [91, 123, 3, -73, 8, -102, -19, 118, 88, 58])
    np.testing.assert_array_equal(oracle.XorShiftRng(1234).u8(10), [91, 123, 3, 183, 8, 154, 237, 118, 88, 58])
    np.testing.assert_array_equal(
        oracle.XorShiftRng(1234).i32(10),
        [-533893029, -1874043781, -2014135805, -1501708361, 330844424, 1872264090, -1812926995, -306325642,
         692957528, -1439925190])


def test_reduced_range_rng(oracle):
    # rten-gemm/src/reduced_range_rng.rs:37-57: i7 / u7 ranges
    v = oracle.XorShiftRng(1234).i8(4096, reduce_range=True)
    assert v.min() >= -64 and v.max() <= 63
    u = oracle.XorShiftRng(1234).u8(4096, reduce_range=True)
    assert u.max() <= 127


# ---- GEMM: rten-gemm/src/tests.rs:249-268 and the sweeps ---------------------------------
def ref_gemm(a, b, alpha=1.0, beta=0.0, c=None, bias=None, kind=None, a_zp=None, b_zp=None):
    """rten-gemm/src/tests.rs:90-133 reference_gemm (float64 accumulate here for f32)."""
    if np.issubdtype(a.dtype, np.integer):
        az = np.zeros(a.shape[0], np.int64) if a_zp is None else np.asarray(a_zp, np.int64)
        bz = np.zeros(b.shape[1], np.int64) if b_zp is None else np.asarray(b_zp, np.int64)
        acc = (a.astype(np.int64) - az[:, None]) @ (b.astype(np.int64) - bz[None, :])
        return acc.astype(np.int64).astype(np.int32)  # wrap
    out = alpha * (a.astype(np.float64) @ b.astype(np.float64))
    if c is not None and beta != 0:
        out = out + beta * c
    if bias is not None:
        out = out + (bias[None, :] if kind == "row" else bias[:, None])
    return out.astype(np.float32)


def test_simple_gemm(oracle):
    a = np.array([[1, 2], [3, 4]], np.float32)
    b = np.array([[5, 6], [7, 8]], np.float32)
    np.testing.assert_array_equal(oracle.gemm_f32(a, b), [[19, 22], [43, 50]])
    np.testing.assert_array_equal(oracle.gemm_u8i8(a.astype(np.uint8), b.astype(np.int8)), [[19, 22], [43, 50]])


@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (2, 2, 2), (5, 7, 10), (6, 32, 256), (7, 33, 257), (64, 128, 300),
                                   (65, 1025, 20), (1, 130, 17), (1, 8, 600), (10, 5, 0)])
def test_gemm_f32_sizes(oracle, m, n, k):
    # size sweep straddling MR/NR/kc like rten-gemm/src/tests.rs:336-362
    r = oracle.XorShiftRng(1234)
    a = r.f32((m, k))
    b = r.f32((k, n))
    got = oracle.gemm_f32(a, b)
    assert oracle.expect_equal(got, ref_gemm(a, b), atol=1e-6 * max(k, 1), rtol=1e-5)


def test_gemm_f32_strided_b(oracle):
    # transposed / strided B: rten-gemm/src/tests.rs:522-568
    r = oracle.XorShiftRng(1234)
    a = r.f32((9, 40))
    bt = r.f32((21, 40))
    got = oracle.gemm_f32(a, bt.T)
    assert oracle.expect_equal(got, ref_gemm(a, bt.T), atol=1e-5)
    big = r.f32((40, 50))
    got = oracle.gemm_f32(a, big[:, ::2])
    assert oracle.expect_equal(got, ref_gemm(a, big[:, ::2]), atol=1e-5)


@pytest.mark.parametrize("alpha", [0.0, 0.5, 1.0, 2.0])
@pytest.mark.parametrize("beta", [0.0, 0.5, 1.0, 2.0])
def test_gemm_alpha_beta(oracle, alpha, beta):
    # rten-gemm/src/tests.rs:570-674 incl. NaN-poisoned output with beta = 0
    r = oracle.XorShiftRng(1234)
    a = r.f32((10, 300))
    b = r.f32((300, 15))
    c = r.f32((10, 15))
    if beta == 0.0:
        c = np.full((10, 15), np.nan, np.float32)
    got = oracle.gemm_f32(a, b, c=c, alpha=alpha, beta=beta)
    assert not np.isnan(got).any()
    assert oracle.expect_equal(got, ref_gemm(a, b, alpha, beta, None if beta == 0 else c), atol=1e-4)


@pytest.mark.parametrize("kind", ["row", "column"])
@pytest.mark.parametrize("m,n,k", [(10, 15, 300), (1, 15, 20), (3, 1, 7)])
def test_gemm_bias(oracle, kind, m, n, k):
    # rten-gemm/src/tests.rs:676-718
    r = oracle.XorShiftRng(1234)
    a = r.f32((m, k))
    b = r.f32((k, n))
    bias = r.f32((n if kind == "row" else m,))
    got = oracle.gemm_f32(a, b, bias=bias, bias_kind=kind)
    assert oracle.expect_equal(got, ref_gemm(a, b, bias=bias, kind=kind), atol=1e-4)


@pytest.mark.parametrize("m,n,k", [(5, 7, 10), (1, 5, 10), (1, 8, 4), (1, 16, 4), (1, 1, 2), (1, 256, 10)])
def test_gemm_u8i8_zero_points(oracle, m, n, k):
    # rten-gemm/src/tests.rs:431-481: za[i] = i, zb[j] = j
    r = oracle.XorShiftRng(1234)
    a = r.u8((m, k))
    b = r.i8((k, n))
    za = np.arange(m).astype(np.uint8)
    zb = np.arange(n).astype(np.int8)
    for azp, bzp in [(None, None), (za, None), (None, zb), (za, zb)]:
        got = oracle.gemm_u8i8(a, b, azp, bzp)
        np.testing.assert_array_equal(got, ref_gemm(a, b, a_zp=azp, b_zp=bzp))


def test_gemm_errors(oracle):
    # rten-gemm/src/tests.rs:280-292,503-519
    z = np.zeros
    with pytest.raises(oracle.OpError, match="KSizeMismatch"):
        oracle.gemm_f32(z((2, 3), np.float32), z((4, 2), np.float32))
    with pytest.raises(oracle.OpError, match="WrongBiasSize"):
        oracle.gemm_f32(z((2, 3), np.float32), z((3, 2), np.float32), bias=z(5, np.float32), bias_kind="row")
    with pytest.raises(oracle.OpError, match="WrongQuantParamSize"):
        oracle.gemm_u8i8(z((2, 3), np.uint8), z((3, 2), np.int8), a_zp=z(3, np.uint8))


# ---- MatMul / Gemm ops: src/ops/matmul.rs ------------------------------------------------
def test_gemm_op(oracle):
    r = oracle.XorShiftRng(1234)
    a = r.f32((3, 10))
    b = r.f32((10, 8))
    c = r.f32((8,))
    np.testing.assert_allclose(oracle.gemm_op(a, b, c, 1.0, 1.0), a @ b + c, rtol=1e-5)
    np.testing.assert_allclose(oracle.gemm_op(a.T.copy(), b.T.copy(), None, 0.5, 1.0, True, True), 0.5 * (a @ b), rtol=1e-5)
    with pytest.raises(oracle.OpError, match="Columns of first matrix does not match rows of second matrix"):
        oracle.gemm_op(a, b.T.copy())
    with pytest.raises(oracle.OpError, match="Cannot broadcast c to output shape"):
        oracle.gemm_op(a, b, np.zeros((5,), np.float32))


@pytest.mark.parametrize("ashape,bshape", [
    ((3, 10), (10, 8)), ((2, 3, 10), (10, 8)), ((3, 10), (2, 10, 8)), ((2, 3, 10), (2, 10, 8)),
    ((2, 1, 3, 10), (1, 4, 10, 8)), ((10,), (10, 8)), ((3, 10), (10,)), ((10,), (10,)),
    ((2, 0, 10), (10, 8)), ((3, 10), (10, 0))])
def test_matmul_shapes(oracle, ashape, bshape):
    # shape table src/ops/matmul.rs:1104-1165, zero-sized dims :1344-1361
    r = oracle.XorShiftRng(1234)
    a = r.f32(ashape)
    b = r.f32(bshape)
    got = oracle.matmul(a, b)
    exp = np.matmul(a.astype(np.float64), b.astype(np.float64))
    assert got.shape == exp.shape
    np.testing.assert_allclose(got, exp, rtol=1e-5, atol=1e-6)


def test_fused_matmul(oracle):
    r = oracle.XorShiftRng(1234)
    a = r.f32((2, 5, 12))
    b = r.f32((12, 7))
    bias = r.f32((7,))
    got = oracle.matmul(a, b, bias=bias, alpha=0.125)
    np.testing.assert_allclose(got, 0.125 * (a @ b) + bias, rtol=1e-5, atol=1e-6)


def test_matmul_errors(oracle):
    # src/ops/matmul.rs:1292-1315
    z = lambda *s: np.zeros(s, np.float32)
    with pytest.raises(oracle.OpError, match="Columns of first matrix does not match rows of second matrix"):
        oracle.matmul(z(1, 2), z(3, 1))
    with pytest.raises(oracle.OpError, match="Inputs must have >= 1 dimensions"):
        oracle.matmul(np.float32(1.0), z(3, 1))
    with pytest.raises(oracle.OpError, match="Cannot broadcast shapes"):
        oracle.matmul(z(2, 2, 2), z(3, 2, 2))


def _ref_mmi(a, b, az, bz):
    """src/ops/matmul.rs:1576-1655 reference_matmul_integer (pure i32)."""
    a = np.asarray(a)
    b = np.asarray(b)
    a_vec, b_vec = a.ndim == 1, b.ndim == 1
    a2 = a[None, :] if a_vec else a
    b2 = b[:, None] if b_vec else b
    azv = np.zeros(a2.shape[-2], np.int64) if az is None else np.broadcast_to(np.asarray(az, np.int64), (a2.shape[-2],))
    bzv = np.zeros(b2.shape[-1], np.int64) if bz is None else np.broadcast_to(np.asarray(bz, np.int64), (b2.shape[-1],))
    out = np.matmul(a2.astype(np.int64) - azv[:, None], b2.astype(np.int64) - bzv[None, :])
    if a_vec:
        out = out.squeeze(-2)
    if b_vec:
        out = out.squeeze(-1)
    return out.astype(np.int32)


def test_matmul_integer_literals(oracle):
    # src/ops/matmul.rs:1375-1500 (u8 x i8 literal cases)
    A = np.array([[1, 2], [3, 4]], np.uint8)
    B = np.array([[5, 6], [7, 8]], np.int8)
    cases = [
        (A, B, None, None),
        (A, B, np.uint8(127), np.int8(-50)),
        (A, B, np.array([1, 2], np.uint8), np.array([3, 4], np.int8)),
        (np.zeros((3, 2, 2), np.uint8), B, np.array([1, 2], np.uint8), np.array([3, 4], np.int8)),
        (np.array([[1, 2, 3, 4]], np.uint8), np.array([[5, 6], [7, 8], [9, 10], [11, 12]], np.int8),
         np.array([1], np.uint8), np.array([3, 4], np.int8)),
        (np.array([1, 2], np.uint8), np.array([[1, 2], [3, 4]], np.int8), np.array([1], np.uint8), np.array([2, 3], np.int8)),
        (A, np.array([1, 2], np.int8), np.array([1, 2], np.uint8), np.array([3], np.int8)),
        (np.zeros((0, 2), np.uint8), np.zeros((2, 3), np.int8), None, None),
    ]
    for a, b, az, bz in cases:
        got = oracle.matmul_integer(a, b, az, bz)
        exp = _ref_mmi(a, b, az, bz)
        assert got.shape == exp.shape
        np.testing.assert_array_equal(got, exp)
    # scalar zero points literal: (a-127)(b+50)
    np.testing.assert_array_equal(oracle.matmul_integer(A, B, np.uint8(127), np.int8(-50)),
                                  (A.astype(int) - 127) @ (B.astype(int) + 50))


def test_matmul_integer_errors(oracle):
    A = np.array([[1, 2], [3, 4]], np.uint8)
    B = np.array([[5, 6], [7, 8]], np.int8)
    with pytest.raises(oracle.OpError, match="Zero point has incorrect size"):
        oracle.matmul_integer(A, B, np.array([1, 2, 4], np.uint8), np.array([3, 4], np.int8))
    with pytest.raises(oracle.OpError, match="Only scalar or vector zero points are supported"):
        oracle.matmul_integer(A, B, np.full((2, 2), 2, np.uint8), None)
    with pytest.raises(oracle.OpError, match="Columns of first matrix does not match rows of second matrix"):
        oracle.matmul_integer(np.zeros((1, 2), np.uint8), np.zeros((3, 1), np.int8))
    with pytest.raises(oracle.OpError, match="Inputs must have >= 1 dimensions"):
        oracle.matmul_integer(np.zeros((), np.uint8), np.zeros((3, 1), np.int8))
    with pytest.raises(oracle.OpError, match="Cannot broadcast shapes"):
        oracle.matmul_integer(np.zeros((2, 2, 2), np.uint8), np.zeros((3, 2, 2), np.int8))


@pytest.mark.parametrize("adt,bdt", [(np.uint8, np.int8), (np.uint8, np.uint8), (np.int8, np.int8), (np.int8, np.uint8)])
def test_matmul_integer_signedness(oracle, adt, bdt):
    # four signedness combos: src/ops/matmul.rs:1657-1750
    r = oracle.XorShiftRng(1234)
    a = r.u8((2, 5, 20)).view(adt) if adt == np.int8 else r.u8((2, 5, 20))
    b = r.u8((20, 9)).view(bdt) if bdt == np.int8 else r.u8((20, 9))
    az = np.arange(5).astype(adt)
    bz = (np.arange(9) - 4).astype(bdt) if bdt == np.int8 else np.arange(9).astype(bdt)
    np.testing.assert_array_equal(oracle.matmul_integer(a, b, az, bz), _ref_mmi(a, b, az, bz))


def test_cast_scale(oracle):
    # src/ops/matmul.rs:959-977
    d = np.array([[1, 2], [3, 4]], np.int32)
    np.testing.assert_array_equal(oracle.cast_scale(d, np.array([2., 3.], np.float32)), [[2., 6.], [6., 12.]])
    np.testing.assert_array_equal(oracle.cast_scale(d, np.float32(2.)), [[2., 4.], [6., 8.]])
    with pytest.raises(oracle.OpError, match="Scale length does not match tensor columns"):
        oracle.cast_scale(d, np.array([2., 3., 4.], np.float32))


# ---- Conv: src/ops/conv.rs ---------------------------------------------------------------
KERNEL = np.array([0.3230, 0.7632, 0.4616, 0.8837, 0.5898, 0.3424, 0.2101, 0.7821, 0.6861], np.float32).reshape(1, 1, 3, 3)
INPUT = np.array([0.5946, 0.8249, 0.0448, 0.9552, 0.2041, 0.2501, 0.2693, 0.1007, 0.8862], np.float32).reshape(1, 1, 3, 3)


def test_conv_torch_goldens(oracle):
    # src/ops/conv.rs:783-839 (PyTorch-derived)
    same = np.array([1.5202, 1.5592, 0.9939, 1.7475, 2.6358, 1.3428, 1.0165, 1.1806, 0.8685], np.float32).reshape(1, 1, 3, 3)
    assert eq_1e4(oracle.conv(INPUT, KERNEL, None, [1, 1, 1, 1]), same)
    assert eq_1e4(oracle.conv(INPUT, KERNEL, None, [0, 0, 0, 0]), [[[[2.6358]]]])
    assert eq_1e4(oracle.conv(INPUT, KERNEL, np.array([1.0], np.float32), [0, 0, 0, 0]), [[[[3.6358]]]])
    assert eq_1e4(oracle.conv(INPUT, KERNEL, None, "same"), same)


def test_conv_depthwise_golden(oracle):
    # src/ops/conv.rs:990-1030 (test_conv_depthwise): one input channel per output channel, groups = 3
    x = np.array([0.5946, 0.8249, 0.0448, 0.9552, 0.2041, 0.2501, 0.2693, 0.1007, 1.5202, 1.5592, 0.9939, 1.7475], np.float32).reshape(1, 3, 2, 2)
    w = np.array([-0.0862, -0.4111, 0.0813, 0.4993, -0.4641, 0.1715, -0.0532, -0.2429, -0.4325, 0.4273, 0.4180, 0.4338], np.float32).reshape(3, 1, 2, 2)
    bias = np.array([0.1, 0.2, 0.3], np.float32)
    want = np.array([0.09020272 + 0.1, -0.09061745 + 0.2, 1.1822754 + 0.3], np.float32).reshape(1, 3, 1, 1)
    assert eq_1e4(oracle.conv(x, w, bias, [0, 0, 0, 0], groups=3), want)


def ref_conv(x, w, bias, pads, groups, strides, dil, x_zp=None, w_zp=None):
    """src/ops/conv.rs:629-747 reference_conv: 7-deep loop; padded taps are SKIPPED."""
    integer = np.issubdtype(x.dtype, np.integer)
    B, C, H, W = x.shape
    O, cg, kh, kw = w.shape
    pt, pl, pb, pr = pads
    oh = (H + pt + pb - dil[0] * (kh - 1) - 1) // strides[0] + 1
    ow = (W + pl + pr - dil[1] * (kw - 1) - 1) // strides[1] + 1
    y = np.zeros((B, O, oh, ow), np.int64 if integer else np.float64)
    og = O // groups
    for n in range(B):
        for o in range(O):
            g = o // og
            for oy in range(oh):
                for ox in range(ow):
                    acc = 0
                    for c in range(cg):
                        for ky in range(kh):
                            for kx in range(kw):
                                iy = oy * strides[0] - pt + ky * dil[0]
                                ix = ox * strides[1] - pl + kx * dil[1]
                                if 0 <= iy < H and 0 <= ix < W:
                                    xv = x[n, g * cg + c, iy, ix]
                                    wv = w[o, c, ky, kx]
                                    if integer:
                                        xv = int(xv) - (0 if x_zp is None else int(x_zp))
                                        wv = int(wv) - (0 if w_zp is None else int(w_zp[o]))
                                    else:
                                        xv, wv = float(xv), float(wv)
                                    acc += xv * wv
                    y[n, o, oy, ox] = acc + (0 if bias is None else float(bias[o]))
    return y.astype(np.int32 if integer else np.float32)


@pytest.mark.parametrize("case", [
    dict(x=(1, 3, 8, 8), w=(4, 3, 3, 3), pads=[1, 1, 1, 1], strides=(1, 1), dil=(1, 1), groups=1),
    dict(x=(2, 4, 9, 7), w=(6, 2, 3, 2), pads=[0, 1, 2, 0], strides=(2, 1), dil=(1, 1), groups=2),
    dict(x=(1, 2, 10, 10), w=(3, 2, 3, 3), pads=[2, 2, 2, 2], strides=(2, 3), dil=(2, 2), groups=1),
    dict(x=(2, 5, 6, 6), w=(7, 5, 1, 1), pads=[0, 0, 0, 0], strides=(1, 1), dil=(1, 1), groups=1),
    dict(x=(1, 4, 5, 5), w=(4, 1, 3, 3), pads=[1, 1, 1, 1], strides=(1, 1), dil=(1, 1), groups=4),
])
def test_conv_vs_reference_conv(oracle, case):
    # stride / dilation / padding / groups sweeps vs reference_conv: src/ops/conv.rs:1131-1319
    r = oracle.XorShiftRng(1234)
    x = r.f32(case["x"])
    w = r.f32(case["w"])
    bias = r.f32((case["w"][0],))
    got = oracle.conv(x, w, bias, case["pads"], case["groups"], case["strides"], case["dil"])
    exp = ref_conv(x, w, bias, case["pads"], case["groups"], case["strides"], case["dil"])
    assert oracle.expect_equal(got, exp, atol=1e-5)


def test_conv_1d(oracle):
    r = oracle.XorShiftRng(1234)
    x = r.f32((2, 3, 11))
    w = r.f32((4, 3, 3))
    got = oracle.conv(x, w, None, [1, 1], 1, (2,), (1,))
    exp = ref_conv(x[:, :, None, :], w[:, :, None, :], None, [0, 1, 0, 1], 1, (1, 2), (1, 1))
    assert oracle.expect_equal(got, exp[:, :, 0, :], atol=1e-5)


def test_conv_errors(oracle):
    # src/ops/conv.rs:1194-1252
    z = lambda *s: np.zeros(s, np.float32)
    with pytest.raises(oracle.OpError, match=r"Input channels \(per group\) does not match kernel input channels"):
        oracle.conv(z(1, 3, 5, 5), z(2, 2, 3, 3))
    with pytest.raises(oracle.OpError, match="Group count must be > 0"):
        oracle.conv(z(1, 2, 5, 5), z(2, 2, 3, 3), groups=0)
    with pytest.raises(oracle.OpError, match="Input channel count not divisible by groups"):
        oracle.conv(z(1, 3, 5, 5), z(2, 1, 3, 3), groups=2)
    with pytest.raises(oracle.OpError, match="Output channel count not divisible by groups"):
        oracle.conv(z(1, 4, 5, 5), z(3, 2, 3, 3), groups=2)
    with pytest.raises(oracle.OpError, match="Input too small for kernel size"):
        oracle.conv(z(1, 1, 2, 2), z(1, 1, 3, 3))
    with pytest.raises(oracle.OpError, match="Strides must be > 0"):
        oracle.conv(z(1, 1, 5, 5), z(1, 1, 3, 3), strides=(0, 1))
    with pytest.raises(oracle.OpError, match="expected 2 stride values"):
        oracle.conv(z(1, 1, 5, 5), z(1, 1, 3, 3), strides=(1,))


@pytest.mark.parametrize("xdt,wdt", [(np.uint8, np.int8), (np.uint8, np.uint8), (np.int8, np.int8), (np.int8, np.uint8)])
def test_conv_integer_cases(oracle, xdt, wdt):
    # src/ops/conv.rs:1429-1497: x_zp=12, w_zp=[1,2,3(,4)], zero padding only (as the reference tests)
    rng = oracle.XorShiftRng(1234)
    krng = oracle.XorShiftRng(5678)
    mk = lambda r, s, dt: (r.u8(s).view(np.int8) if dt == np.int8 else r.u8(s))
    cases = [((1, 2, 5, 5), (1, 2, 3, 3), 12, [1], 1), ((1, 2, 5, 5), (3, 2, 3, 3), 12, [1, 2, 3], 1),
             ((1, 4, 5, 5), (4, 2, 3, 3), 12, [1, 2, 3, 4], 2), ((1, 2, 5, 5), (1, 2, 3, 3), None, None, 1),
             ((1, 2, 5, 5), (1, 2, 1, 1), 12, [1], 1), ((1, 2, 1, 1), (1, 2, 1, 1), 12, [1], 1)]
    for xs, ws, xz, wz, g in cases:
        x = mk(rng, xs, xdt)
        w = mk(krng, ws, wdt)
        xzp = None if xz is None else np.array(xz, xdt)
        wzp = None if wz is None else np.array(wz, wdt)
        got = oracle.conv_integer(x, w, xzp, wzp, groups=g)
        exp = ref_conv(x, w, None, [0, 0, 0, 0], g, (1, 1), (1, 1), xz, wz)
        np.testing.assert_array_equal(got, exp)


def test_conv_integer_padding_production_path(oracle):
    # G3 (unpinned by the reference's tests): padded taps behave as literal 0 in the shifted-i8 domain,
    # i.e. as the value 128 for u8 inputs and 0 for i8 inputs, and still receive the -x_zp correction.
    rng = oracle.XorShiftRng(1234)
    x = rng.u8((1, 2, 5, 5))
    w = oracle.XorShiftRng(5678).i8((3, 2, 3, 3))
    got = oracle.conv_integer(x, w, np.uint8(12), np.array([1, 2, 3], np.int8), padding=[1, 1, 1, 1])
    xp = np.full((1, 2, 7, 7), 128, np.uint8)
    xp[:, :, 1:6, 1:6] = x
    exp = ref_conv(xp, w, None, [0, 0, 0, 0], 1, (1, 1), (1, 1), 12, [1, 2, 3])
    np.testing.assert_array_equal(got, exp)


def test_conv_integer_to_float(oracle):
    # src/ops/conv.rs:1571-1587
    rng = oracle.XorShiftRng(1234)
    x = rng.u8((1, 2, 5, 5))
    w = oracle.XorShiftRng(5678).i8((3, 2, 3, 3))
    i = oracle.conv_integer(x, w, np.uint8(12), np.array([1, 2, 3], np.int8))
    f = oracle.conv_integer_to_float(x, w, np.uint8(12), np.array([1, 2, 3], np.int8), np.float32(0.1))
    np.testing.assert_array_equal(f, i.astype(np.float32) * np.float32(0.1))
    f1 = oracle.conv_integer_to_float(x, w, np.uint8(12), np.array([1, 2, 3], np.int8), np.array([0.1], np.float32))
    np.testing.assert_array_equal(f, f1)
    with pytest.raises(oracle.OpError, match="scale should be a scalar"):
        oracle.conv_integer_to_float(x, w, np.uint8(12), np.array([1, 2, 3], np.int8), np.array([0.1, 0.2, 0.3], np.float32))


# ---- LayerNormalization: src/ops/norm.rs:1088-1167 ---------------------------------------
LN_IN = np.array([[[0.9562, 0.0572], [0.4366, 0.5655], [0.2017, 0.0230], [0.7941, 0.1554], [0.3226, 0.120]]], np.float32)


def test_layer_norm_goldens(oracle):
    got = oracle.layer_norm(LN_IN, np.array([0.0751, 0.6952], np.float32), np.array([0.9993, 0.7632], np.float32), -1)
    assert eq_1e4(got, [[[1.0744, 0.0680], [0.9243, 1.4576], [1.0744, 0.0684], [1.0744, 0.0680], [1.0744, 0.0683]]])
    got = oracle.layer_norm(LN_IN, np.full((5, 2), 1.1, np.float32), np.full((5, 2), 0.1, np.float32), -2)
    assert eq_1e4(got, [[[2.2467697, -1.0079411], [0.36562642, 0.83229196], [-0.48479798, -1.1317577],
                         [1.6599079, -0.65242106], [-0.04709549, -0.7805821]]])
    x = np.array([[0., 1., 2., 3.]], np.float32)
    assert eq_1e4(oracle.layer_norm(x, np.float32(2.0), np.float32(0.5), -1), [[-2.1833, -0.3944, 1.3944, 3.1833]])
    assert eq_1e4(oracle.layer_norm(x, np.float32(2.0), None, -1), [[-2.6833, -0.8944, 0.8944, 2.6833]])
    with pytest.raises(oracle.OpError, match="`scale` is not broadcastable to normalized axes of input"):
        oracle.layer_norm(np.ones((2, 3), np.float32), np.ones((2, 3), np.float32), None, -1)
    with pytest.raises(oracle.OpError, match="`bias` is not broadcastable to normalized axes of input"):
        oracle.layer_norm(np.ones((2, 3), np.float32), np.ones(3, np.float32), np.ones((2, 3), np.float32), -1)


# ---- Softmax: src/ops/norm.rs:1302-1411, rten-vecmath/src/softmax.rs:265-285 -------------
def test_softmax_goldens(oracle):
    x = np.array([0.1634, 0.8647, 0.6401, 0.8265, 0.0560, 0.2304], np.float32)
    assert eq_1e4(oracle.softmax(x, 0), [0.1172, 0.2362, 0.1887, 0.2274, 0.1052, 0.1253])
    assert oracle.softmax(np.zeros((0,), np.float32), 0).shape == (0,)
    x2 = x.reshape(2, 3)
    assert eq_1e4(oracle.softmax(x2, 1), [[0.2161, 0.4358, 0.3481], [0.4966, 0.2298, 0.2736]])
    assert eq_1e4(oracle.softmax(x2, 0), [[0.3400, 0.6918, 0.6010], [0.6600, 0.3082, 0.3990]])
    m = np.array([[0.1634, 0.8647, 0.6401, 0.8265, 0.0560]] * 2, np.float32)
    assert eq_1e4(oracle.softmax(m, 1), [[0.1339, 0.2701, 0.2157, 0.2599, 0.1203]] * 2)


def test_softmax_vecmath_1ulp(oracle):
    # rten-vecmath/src/softmax.rs:265-272 (1 ULP vs the listed values)
    x = np.array([0.1634, 0.8647, 0.6401, 0.8265, 0.0560, 0.2304], np.float32)
    exp = np.array([0.11715934, 0.23623686, 0.18871443, 0.2273828, 0.10522857, 0.12527795], np.float32)
    got = oracle.softmax(x, 0)
    ulp = np.abs(got.view(np.int32).astype(np.int64) - exp.view(np.int32).astype(np.int64))
    assert ulp.max() <= 1


def test_softmax_transposed(oracle):
    # src/ops/norm.rs:1343-1369
    x = np.array([0.6427, 0.7435, 0.9762, 0.0611, 0.1249, 0.9742, 0.5826, 0.4704, 0.1420, 0.8376, 0.6692, 0.7090,
                  0.2448, 0.9083, 0.2881, 0.4971], np.float32).reshape(4, 4)
    exp = np.array([0.3480, 0.2073, 0.2109, 0.2337, 0.2204, 0.2776, 0.2421, 0.2599, 0.3433, 0.2316, 0.2525, 0.1725,
                    0.1677, 0.2525, 0.3205, 0.2593], np.float32).reshape(4, 4)
    assert eq_1e4(oracle.softmax(x.T, 1), exp)


def test_softmax_properties(oracle):
    # lane-sum ~ 1 (:1374-1394), NaN flush (:1398-1411)
    x = oracle.XorShiftRng(1234).f32((4, 512))
    assert np.all(np.abs(oracle.softmax(x, 1).sum(1) - 1.0) < 1e-3)
    ninf = np.full(3, -np.inf, np.float32)
    assert np.isnan(oracle.softmax(ninf, 0)).all()
    np.testing.assert_array_equal(oracle.softmax(ninf, 0, flush_nans_to_zero=True), [0., 0., 0.])


def test_add_softmax(oracle):
    # src/ops/attention.rs:30-121
    r = oracle.XorShiftRng(1234)
    qk = r.f32((2, 3, 4, 8))
    m = r.f32((2, 1, 1, 8))
    got = oracle.add_softmax(qk, m)
    np.testing.assert_array_equal(got, oracle.softmax(qk + m, -1))
    np.testing.assert_array_equal(oracle.add_softmax(m, qk), got)
    with pytest.raises(oracle.OpError, match="Cannot broadcast inputs"):
        oracle.add_softmax(qk, np.zeros((3, 8), np.float32)[:, :5])


# ---- Erf / Gelu / Tanh / Exp -------------------------------------------------------------
def test_erf_goldens(oracle):
    # src/ops/unary_elementwise.rs:989-1019
    got = oracle.erf(np.array([-2.0, -0.5, 0.5, 2.0], np.float32))
    exp = np.array([-0.9953222650189527, -0.5204998778130465, 0.5204998778130465, 0.9953222650189527], np.float32)
    assert oracle.expect_equal(got, exp)
    sp = oracle.erf(np.array([np.nan, 0., np.inf, -np.inf], np.float32))
    assert np.isnan(sp[0]) and sp[1] == 0 and sp[2] == 1 and sp[3] == -1


def test_erf_gelu_accuracy(oracle):
    # rten-vecmath/src/erf.rs:126 (6.631017e-7 max abs err vs libm), :170 (approx gelu 5e-7... vs tanh formula)
    import math
    x = np.arange(-6, 6, 0.001, dtype=np.float32)
    true = np.array([math.erf(float(v)) for v in x])
    assert np.max(np.abs(oracle.erf(x) - true)) <= 7e-7
    g = 0.5 * x.astype(np.float64) * (1 + true_erf_scaled(x))
    assert np.max(np.abs(oracle.gelu(x) - g)) <= 2e-6
    ag = 0.5 * x.astype(np.float64) * (1 + np.tanh(np.sqrt(2 / np.pi) * (x.astype(np.float64) + 0.044715 * x.astype(np.float64) ** 3)))
    assert np.max(np.abs(oracle.gelu(x, approximate=True) - ag)) <= 2e-6


def true_erf_scaled(x):
    import math
    return np.array([math.erf(float(v) / math.sqrt(2.0)) for v in x])


def test_exp_tanh_ulps(oracle):
    # rten-vecmath/src/exp.rs:284 (<= 1 ULP vs f32::exp), tanh.rs:76 (<= 3 ULP)
    x = np.arange(-6, 6, 0.001, dtype=np.float32)

    def ulps(a, b):
        return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64)).max()

    assert ulps(oracle.exp(x), np.exp(x.astype(np.float64)).astype(np.float32)) <= 1
    assert ulps(oracle.tanh(x), np.tanh(x.astype(np.float64)).astype(np.float32)) <= 3
    assert oracle.exp(np.array([104.0, -104.0, 0.0], np.float32)).tolist() == [np.inf, 0.0, 1.0]


# ---- Quantisation: src/ops/quantize.rs:704-768, rten-vecmath/src/quantize.rs:86-124 ------
def test_dynamic_quantize_linear(oracle):
    r = oracle.XorShiftRng(1234)
    x = (r.f32((5, 1000)) - np.float32(0.3)) * np.float32(4.0)
    y, scale, zp = oracle.dynamic_quantize_linear(x)
    assert y.dtype == np.uint8 and y.shape == x.shape
    # ONNX definition
    lo, hi = min(x.min(), 0), max(x.max(), 0)
    s = np.float32((np.float32(hi) - np.float32(lo)) / np.float32(255))
    assert scale == s
    z = np.clip(np.round(np.float32(0) - np.float32(lo) / s), 0, 255)
    assert zp == z
    deq = (y.astype(np.float32) - np.float32(zp)) * scale
    assert np.max(np.abs(deq - x)) <= scale * 0.5 + 1e-6      # round-trip bound (quantize.rs:704-768)
    # scalar reference: reference_quantize (rten-vecmath/src/quantize.rs:90-97)
    inv = np.float32(1.0) / s
    ref = np.clip(np.rint(x * inv) + np.float32(zp), 0, 255).astype(np.uint8)
    np.testing.assert_array_equal(y, ref)
    # empty input
    y, scale, zp = oracle.dynamic_quantize_linear(np.zeros((0, 3), np.float32))
    assert scale == 1.0 and zp == 0 and y.shape == (0, 3)


def test_quantize_u8_vecmath_case(oracle):
    src = oracle.XorShiftRng(1234).f32(65)
    got = oracle.quantize_u8(src, 5.2, 10)
    ref = np.clip(np.rint(src * np.float32(5.2)) + 10, 0, 255).astype(np.uint8)
    np.testing.assert_array_equal(got, ref)


# ---- cross-check with PyTorch CPU (independent second opinion) --------------
def test_vs_torch(oracle):
    torch = pytest.importorskip("torch")
    import torch.nn.functional as F
    r = oracle.XorShiftRng(1234)
    x = r.uniform((2, 6, 13, 11))
    w = r.uniform((8, 3, 3, 3))
    b = r.uniform((8,))
    got = oracle.conv(x, w, b, [1, 2, 1, 0], 2, (2, 1), (1, 2))
    xt = F.pad(torch.from_numpy(x), (2, 0, 1, 1))
    exp = F.conv2d(xt, torch.from_numpy(w), torch.from_numpy(b), stride=(2, 1), dilation=(1, 2), groups=2).numpy()
    np.testing.assert_allclose(got, exp, rtol=1e-5, atol=1e-5)
    h = r.uniform((7, 768))
    g, be = r.uniform((768,)), r.uniform((768,))
    np.testing.assert_allclose(oracle.layer_norm(h, g, be, -1, 1e-12),
                               F.layer_norm(torch.from_numpy(h), (768,), torch.from_numpy(g), torch.from_numpy(be), 1e-12).numpy(),
                               rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(oracle.softmax(h, -1), F.softmax(torch.from_numpy(h), -1).numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(oracle.gelu(h), F.gelu(torch.from_numpy(h)).numpy(), rtol=0, atol=2e-6)
    mp = oracle.max_pool(x, (3, 3), [1, 1, 1, 1], (2, 2))
    np.testing.assert_array_equal(mp, F.max_pool2d(torch.from_numpy(x), 3, 2, 1).numpy())
    np.testing.assert_allclose(oracle.global_average_pool(x), F.adaptive_avg_pool2d(torch.from_numpy(x), 1).numpy(), rtol=1e-5)
