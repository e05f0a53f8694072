"""In-kernel pipeline trace (CTA 0) of a few launches: prints per-k-block intervals of producer / MMA and per-tile
epilogue times, to see which hand-off bounds the kernel."""
import ctypes as C
import os

os.environ.setdefault("RTEN_B200_F32_MODE", "tf32")  # these tools measure the single-pass TF32 kernels unless told otherwise
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import rten_b200 as rt  # noqa: E402

torch.cuda.set_device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = rt.Context(0, stream=stream.cuda_stream)


def trace(fn, name):
    fn()
    ctx.sync()
    ctx.check(ctx.lib.rten_b200_debug_trace(ctx.handle, 1, None))
    fn()
    buf = (C.c_int64 * 8192)()
    ctx.check(ctx.lib.rten_b200_debug_trace(ctx.handle, 0, buf))
    t = np.frombuffer(buf, dtype=np.int64).reshape(4, 2048)
    ph = t[3][1024:1032].copy()
    t[3][1024:1032] = 0
    marks = t[3][1100:1105].copy()
    t[3][1100:1105] = 0
    issue_cost = t[3][680:1360]
    commit_cost = t[3][1364:2044]
    t[3][680:] = 0
    prod, mma, e0, e1 = (r[r > 0] for r in t)
    ic, cc = issue_cost[issue_cost > 0], commit_cost[commit_cost > 0]
    if len(ic):
        print(f"   MMA thread: wait-done->4(8) MMAs issued median {np.median(ic):.0f} clk; commit issue median {np.median(cc):.0f} clk")
    t0 = min(prod.min(), mma.min())
    print(f"== {name}: {len(prod)} k-blocks, {len(e0)} tiles on CTA 0; total {(max(e1.max(), mma.max()) - t0)} clk")
    if marks[0] > 0:
        m0 = marks[0]
        print(f"   timeline (clk from kernel entry): set-up done {marks[1] - m0}, predecessor done {marks[2] - m0}, first TMA issued {prod.min() - m0}, "
              f"first operands landed {mma.min() - m0}, first epilogue start {e0.min() - m0 if len(e0) else -1}, last epilogue end {e1.max() - m0 if len(e1) else -1}, "
              f"thread 0 done {marks[3] - m0}, exit {marks[4] - m0}")
    if len(prod) > 1:
        d = np.diff(prod)
        print(f"   producer slot-acquire interval: median {np.median(d):.0f} mean {d.mean():.0f} max {d.max()}  first 12: {d[:12].tolist()}")
    if len(mma) > 1:
        d = np.diff(mma)
        print(f"   MMA operands-landed interval:   median {np.median(d):.0f} mean {d.mean():.0f} max {d.max()}  first 12: {d[:12].tolist()}")
        n = min(len(prod), len(mma))
        lat = mma[:n] - prod[:n]
        print(f"   TMA issue->landed latency:      median {np.median(lat):.0f} min {lat.min()} max {lat.max()}  first 12: {lat[:12].tolist()}")
    if ph[7] > 0:
        names = ["tmem_ld", "math(+bias/res)", "st.shared(+slow path)", "wait prev store read", "fence.proxy.async", "bar.sync", "store issue"]
        print("   epilogue phases (warp 4, clk per 32-col chunk): " + ", ".join(f"{n} {ph[i] / ph[7]:.0f}" for i, n in enumerate(names)))
    if len(e0):
        print(f"   epilogue duration per tile:     median {np.median(e1 - e0):.0f}  first 6: {(e1 - e0)[:6].tolist()};  tile start interval {np.diff(e0)[:6].tolist()}")


B = 32


def conv_case(ci, co, k, s, p, hw):
    x = rt.from_torch(ctx, torch.randn(B, hw, hw, ci, device="cuda")).permute(0, 3, 1, 2)
    wt = ctx.to_device(np.random.randn(co, ci, k, k).astype(np.float32))
    bias = ctx.to_device(np.zeros(co, np.float32))
    op = rt.Conv(1, (1, 1), (p, p, p, p), (s, s), activation=rt.ACT_RELU)
    pk = op.prepack(ctx, 1, wt)
    y = op.run(ctx, x, wt, bias, packed_w=pk)
    return lambda: op.run(ctx, x, wt, bias, packed_w=pk, out=y)


def gemm_case(m, n, k):
    a = rt.from_torch(ctx, torch.randn(m, k, device="cuda"))
    b = rt.from_torch(ctx, torch.randn(n, k, device="cuda")).permute(1, 0)
    out = ctx.empty((m, n))
    return lambda: rt.MatMul().run(ctx, a, b, out=out)


trace(gemm_case(4096, 4096, 4096), "gemm 4096^3")
trace(conv_case(64, 64, 3, 1, 1, 56), "conv 3x3 64->64 @56")
trace(conv_case(64, 256, 1, 1, 0, 56), "conv 1x1 64->256 @56")
trace(conv_case(256, 64, 1, 1, 0, 56), "conv 1x1 256->64 @56")
trace(conv_case(128, 128, 3, 1, 1, 28), "conv 3x3 128->128 @28")
trace(conv_case(256, 256, 3, 1, 1, 14), "conv 3x3 256->256 @14")
trace(conv_case(1024, 256, 1, 1, 0, 14), "conv 1x1 1024->256 @14")
trace(conv_case(512, 512, 3, 1, 1, 7), "conv 3x3 512->512 @7")
