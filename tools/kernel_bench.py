"""Micro-benchmarks of the tensor-core kernels on one B200 (CUDA events, L2 flushed between launches):
GEMM TFLOP/s (tf32 and int8) and the ResNet-50 / BERT layer shapes.  Also measures the cuBLAS TF32 / bf16
throughput of the same box as a yardstick.  Output: gpurun_out/kernel_bench.json + a table on stdout."""
import json
import os

os.environ.setdefault("RTEN_B200_F32_MODE", "tf32")  # these tools measure the single-pass TF32 kernels unless told otherwise
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import rten_b200 as rt  # noqa: E402


CTX = None


def time_fn(fn, stream, flush, iters=10, warm=3, graph=True):
    """Median / best device time of one call.  Our ops are replayed from a CUDA graph so that the Python/ctypes
    launch path (tens of microseconds) does not sit between the two events."""
    if graph and CTX is not None:
        fn()  # warm the pool
        CTX.graph_begin()
        fn()
        g = CTX.graph_end()
        fn = g.launch
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        fn()
        e.record(stream)
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(min(ts))


def main():
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = rt.Context(0, stream=stream.cuda_stream)
    import os
    ctx.set_autotune(os.environ.get("KB_AUTOTUNE", "1") != "0")
    global CTX
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []

    def rec(name, flops, ms_med, ms_min, extra=None):
        r = {"name": name, "ms_median": ms_med, "ms_min": ms_min, "tflops_median": flops / ms_med / 1e9, "tflops_best": flops / ms_min / 1e9}
        if extra:
            r.update(extra)
        rows.append(r)
        print(f"{name:58s} {ms_med:9.4f} ms  {r['tflops_median']:8.1f} TF/s (best {r['tflops_best']:.1f})", flush=True)

    # ---- yardsticks: cuBLAS on this box
    for n in (4096, 8192):
        a = torch.randn(n, n, device="cuda", dtype=torch.bfloat16)
        b = torch.randn(n, n, device="cuda", dtype=torch.bfloat16)
        rec(f"cublas bf16 {n}^3", 2.0 * n ** 3, *time_fn(lambda: torch.matmul(a, b), stream, flush, graph=False))
        a32, b32 = a.float(), b.float()
        torch.backends.cuda.matmul.allow_tf32 = True
        rec(f"cublas tf32 {n}^3", 2.0 * n ** 3, *time_fn(lambda: torch.matmul(a32, b32), stream, flush, graph=False))
        torch.backends.cuda.matmul.allow_tf32 = False
        rec(f"cublas fp32 {n}^3", 2.0 * n ** 3, *time_fn(lambda: torch.matmul(a32, b32), stream, flush, iters=3, warm=1, graph=False))
        ai = torch.randint(-128, 127, (n, n), device="cuda", dtype=torch.int8)
        bi = torch.randint(-128, 127, (n, n), device="cuda", dtype=torch.int8)
        try:
            rec(f"cublas int8 {n}^3", 2.0 * n ** 3, *time_fn(lambda: torch._int_mm(ai, bi), stream, flush, graph=False))
        except Exception as ex:  # noqa: BLE001
            print("torch._int_mm unavailable:", ex)
        del a, b, a32, b32, ai, bi

    CTX = ctx
    # ---- our GEMM: plain shapes
    mm = rt.MatMul()
    for (m, n, k) in [(8192, 8192, 8192), (4096, 4096, 4096), (2048, 768, 768), (2048, 3072, 768), (2048, 768, 3072), (32, 1000, 2048)]:
        at_ = torch.randn(m, k, device="cuda")
        a = rt.from_torch(ctx, at_)
        bt = torch.randn(n, k, device="cuda")
        torch.backends.cuda.matmul.allow_tf32 = True
        rec(f"cublas tf32 gemm {m}x{n}x{k}", 2.0 * m * n * k, *time_fn(lambda: torch.matmul(at_, bt.T), stream, flush, graph=False))
        torch.backends.cuda.matmul.allow_tf32 = False
        b = rt.from_torch(ctx, bt).permute(1, 0)  # [K, N] view of K-major storage: no packing
        out = ctx.empty((m, n))
        rec(f"ours tf32 gemm {m}x{n}x{k}", 2.0 * m * n * k, *time_fn(lambda: mm.run(ctx, a, b, out=out), stream, flush))
        ai = rt.from_torch(ctx, torch.randint(0, 255, (m, k), device="cuda", dtype=torch.uint8))
        bi = rt.from_torch(ctx, torch.randint(-128, 127, (n, k), device="cuda", dtype=torch.int8)).permute(1, 0)
        outi = ctx.empty((m, n), np.int32)
        mi = rt.MatMulInteger()
        rec(f"ours int8 gemm {m}x{n}x{k}", 2.0 * m * n * k, *time_fn(lambda: mi.run(ctx, ai, bi, out=outi), stream, flush))

    # ---- batched attention shapes (BERT b16): QK^T and PV
    q = rt.from_torch(ctx, torch.randn(16, 128, 768, device="cuda"))
    kk = rt.from_torch(ctx, torch.randn(16, 128, 768, device="cuda"))
    qv = q.view((16, 12, 128, 64), (128 * 768, 64, 768, 1))
    kt = kk.view((16, 12, 64, 128), (128 * 768, 64, 1, 768))
    sc = ctx.empty((16, 12, 128, 128))
    fm = rt.FusedMatMul(0.125)
    rec("ours tf32 QK^T 192x(128x128x64)", 2.0 * 192 * 128 * 128 * 64, *time_fn(lambda: fm.run(ctx, qv, kt, out=sc), stream, flush))
    vv = kk.view((16, 12, 128, 64), (128 * 768, 64, 768, 1))
    att = ctx.empty((16 * 128, 768))
    attv = att.view((16, 12, 128, 64), (128 * 768, 64, 768, 1))
    rec("ours tf32 PV 192x(128x64x128)", 2.0 * 192 * 128 * 64 * 128, *time_fn(lambda: mm.run(ctx, sc, vv, out=attv), stream, flush))

    # ---- ResNet-50 conv layers, batch 32, channels-last
    B = 32
    layers = [("stem 7x7s2 3->64 @224", 3, 64, 7, 2, 3, 224), ("1x1 64->64 @56", 64, 64, 1, 1, 0, 56), ("3x3 64->64 @56", 64, 64, 3, 1, 1, 56),
              ("1x1 64->256 @56", 64, 256, 1, 1, 0, 56), ("1x1 256->64 @56", 256, 64, 1, 1, 0, 56), ("3x3s2 128->128 @56", 128, 128, 3, 2, 1, 56),
              ("3x3 128->128 @28", 128, 128, 3, 1, 1, 28), ("1x1 128->512 @28", 128, 512, 1, 1, 0, 28), ("1x1 512->128 @28", 512, 128, 1, 1, 0, 28),
              ("1x1s2 256->512 @56", 256, 512, 1, 2, 0, 56), ("3x3 256->256 @14", 256, 256, 3, 1, 1, 14), ("1x1 256->1024 @14", 256, 1024, 1, 1, 0, 14),
              ("1x1 1024->256 @14", 1024, 256, 1, 1, 0, 14), ("3x3 512->512 @7", 512, 512, 3, 1, 1, 7), ("1x1 512->2048 @7", 512, 2048, 1, 1, 0, 7),
              ("1x1 2048->512 @7", 2048, 512, 1, 1, 0, 7)]
    for name, ci, co, k, s, p, hw in layers:
        x = rt.from_torch(ctx, torch.randn(B, hw, hw, ci, device="cuda")).permute(0, 3, 1, 2)
        w = ctx.to_device(np.random.randn(co, ci, k, k).astype(np.float32))
        bias = ctx.to_device(np.zeros(co, np.float32))
        op = rt.Conv(1, (1, 1), (p, p, p, p), (s, s), activation=rt.ACT_RELU)
        pk = op.prepack(ctx, 1, w)
        y = op.run(ctx, x, w, bias, packed_w=pk)
        oh = y.shape[2]
        fl = 2.0 * B * co * oh * oh * ci * k * k
        rec(f"ours tf32 conv {name}", fl, *time_fn(lambda: op.run(ctx, x, w, bias, packed_w=pk, out=y), stream, flush))
        if ci % 16 == 0:
            xi = rt.from_torch(ctx, torch.randint(0, 255, (B, hw, hw, ci), device="cuda", dtype=torch.uint8)).permute(0, 3, 1, 2)
            wi = ctx.to_device(np.random.randint(-128, 127, (co, ci, k, k)).astype(np.int8))
            opi = rt.ConvInteger(1, (1, 1), (p, p, p, p), (s, s))
            pki = opi.prepack(ctx, 1, wi)
            zp = ctx.to_device(np.array(128, np.uint8))
            yi = opi.run(ctx, xi, wi, zp, None, packed_w=pki)
            rec(f"ours int8 conv {name}", fl, *time_fn(lambda: opi.run(ctx, xi, wi, zp, None, packed_w=pki, out=yi), stream, flush))

    # ---- HBM-bound row kernels (BERT shapes): GB/s of algorithmic bytes
    def rec_bw(name, nbytes, ms_med, ms_min):
        r = {"name": name, "ms_median": ms_med, "ms_min": ms_min, "gbs_median": nbytes / ms_med / 1e6, "gbs_best": nbytes / ms_min / 1e6}
        rows.append(r)
        print(f"{name:58s} {ms_med:9.4f} ms  {r['gbs_median']:8.1f} GB/s (best {r['gbs_best']:.1f})", flush=True)

    h = rt.from_torch(ctx, torch.randn(2048, 3072, device="cuda"))
    ho = ctx.empty((2048, 3072))
    rec_bw("gelu [2048,3072]", 2 * 4 * 2048 * 3072, *time_fn(lambda: rt.Gelu().run(ctx, h, in_place=True), stream, flush))
    x = rt.from_torch(ctx, torch.randn(2048, 768, device="cuda"))
    g = ctx.to_device(np.ones(768, np.float32))
    b = ctx.to_device(np.zeros(768, np.float32))
    xo = ctx.empty((2048, 768))
    ln = rt.LayerNormalization(-1, 1e-12)
    rec_bw("layernorm [2048,768]", 2 * 4 * 2048 * 768, *time_fn(lambda: ln.run(ctx, x, g, b, out=xo), stream, flush))
    s4 = rt.from_torch(ctx, torch.randn(16, 12, 128, 128, device="cuda"))
    m = ctx.to_device(np.zeros((16, 1, 1, 128), np.float32))
    rec_bw("addsoftmax [16,12,128,128]", 2 * 4 * 16 * 12 * 128 * 128, *time_fn(lambda: rt.AddSoftmax().run(ctx, s4, m, in_place=True), stream, flush))
    big = rt.from_torch(ctx, torch.randn(64 << 20, device="cuda"))
    rec_bw("relu 256 MiB in place", 2 * 4 * (64 << 20), *time_fn(lambda: rt.Relu().run(ctx, big, in_place=True), stream, flush))
    q8 = rt.from_torch(ctx, torch.randn(4096, 768, device="cuda"))
    rec_bw("dynamic_quantize_linear [4096,768]", (4 + 4 + 1) * 4096 * 768, *time_fn(lambda: rt.DynamicQuantizeLinear().run(ctx, q8), stream, flush))

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "kernel_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
