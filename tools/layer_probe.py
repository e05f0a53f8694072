"""Per-layer decomposition of the tcgen05 conv kernel on ResNet-50's layer shapes (batch 32, single-pass TF32):
 * hot time: 20 chained launches of one layer in a CUDA graph (programmatic dependent launch, as in the model graph);
 * K scaling of the 1x1 layers (fixed cost + epilogue = intercept, main loop = slope);
 * in-kernel trace of CTA 0 with the specialised epilogue (RTEN_B200_TRACE_FAST): clocks per k-block and per tile epilogue.
Output: gpurun_out/layer_probe.txt"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RTEN_B200_F32_MODE", "tf32")
os.environ["RTEN_B200_TRACE_FAST"] = "1"

import numpy as np  # noqa: E402
import torch  # noqa: E402

import rten_b200 as rt  # noqa: E402

torch.cuda.set_device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = rt.Context(0, stream=stream.cuda_stream)
out = open(os.path.join(ROOT, "gpurun_out", "layer_probe.txt"), "w")
B = int(os.environ.get("PROBE_BATCH", "32"))
PF, PB = 735e12, 6.56e12


def emit(s):
    print(s, flush=True)
    out.write(s + "\n")
    out.flush()


def make(ci, co, k, s, p, hw, res, act=True):
    x = rt.from_torch(ctx, torch.randn(B, hw, hw, ci, device="cuda")).permute(0, 3, 1, 2)
    w = ctx.to_device((np.random.randn(co, ci, k, k) * 0.05).astype(np.float32))
    bias = ctx.to_device(np.zeros(co, np.float32))
    op = rt.Conv(1, (1, 1), (p, p, p, p), (s, s), activation=rt.ACT_RELU if act else rt.ACT_NONE)
    pk = op.prepack(ctx, 1, w)
    y = op.run(ctx, x, w, bias, packed_w=pk)
    r = None
    if res:
        r = rt.from_torch(ctx, torch.randn(B, y.shape[2], y.shape[3], co, device="cuda")).permute(0, 3, 1, 2)
    run = lambda: op.run(ctx, x, w, bias, packed_w=pk, out=y, residual=r) if res else op.run(ctx, x, w, bias, packed_w=pk, out=y)
    run()
    oh = y.shape[2]
    fl = 2.0 * B * co * oh * oh * ci * k * k
    by = 4.0 * (B * hw * hw * ci + B * oh * oh * co * (2 if res else 1) + co * ci * k * k)
    return run, fl, by


def hot_time(run, reps=20, rounds=5):
    run()
    ctx.graph_begin()
    for _ in range(reps):
        run()
    g = ctx.graph_end()
    g.launch()
    ts = []
    for _ in range(rounds):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        g.launch()
        b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / reps)
    del g
    return float(np.median(ts))


def trace(run):
    run()
    ctx.sync()
    ctx.check(ctx.lib.rten_b200_debug_trace(ctx.handle, 1, None))
    run()
    buf = (C.c_int64 * 8192)()
    ctx.check(ctx.lib.rten_b200_debug_trace(ctx.handle, 0, buf))
    t = np.frombuffer(buf, dtype=np.int64).reshape(4, 2048).copy()
    marks = t[3][1100:1105].copy()
    ph = t[3][1024:1033].copy()
    issue = t[3][1030:1036].copy()
    t[3][680:] = 0
    prod, mma, e0, e1 = (r[r > 0] for r in t)
    if not len(mma) or marks[0] <= 0:
        return "   (no trace)"
    m0 = marks[0]
    n = min(len(e0), len(e1))
    s = (f"   trace CTA0: {len(prod)} k-blocks {n} tiles | setup {marks[1] - m0} pred {marks[2] - m0} firstTMA {prod.min() - m0} landed {mma.min() - m0} "
         f"| k-block interval med {np.median(np.diff(mma)) if len(mma) > 1 else 0:.0f} | epi start {e0.min() - m0 if n else -1} "
         f"epi dur med {np.median(e1[:n] - e0[:n]) if n else 0:.0f} | last epi end {e1.max() - m0 if n else -1} exit {marks[4] - m0}")
    if n > 1:
        s += f" | tile interval {np.diff(e0)[:4].tolist()}"
    if issue[1] > 0:
        s += (f"\n   MMA warp clocks: issue {issue[0]} ({issue[0] / issue[1]:.0f} per instruction, {issue[1]} instructions), waiting for the accumulator {issue[2]}, "
              f"for the patch {issue[3]}, for weight stages {issue[4]}, commits + bookkeeping {issue[5]}")
    if ph[8] > 0 and issue[1] == 0:
        names = ["res-prefetch", "tmem_ld", "res-wait", "math", "st.shared", "wait-prev-store", "fence+bar", "store-issue"]
        s += "\n   epilogue phases (clk per 32-col chunk, warp 4 lane 0): " + ", ".join(f"{nm} {ph[i] / ph[8]:.0f}" for i, nm in enumerate(names)) + f"  ({ph[8]} chunks)"
    return s


def layer_line(name, ci, co, k, s, p, hw, res):
    run, fl, by = make(ci, co, k, s, p, hw, res)
    us = hot_time(run)
    sol = max(fl / PF, by / PB) * 1e6
    emit(f"{name:28s} hot {us:6.1f} us  {fl / us / 1e6:6.1f} TF/s  {by / us / 1e3:6.0f} GB/s  SOL(hbm-cold) {sol:5.1f} us  flop-SOL {fl / PF * 1e6:5.1f} us")
    emit(trace(run))
    return us


LAYERS = [
    ("1x1 64->64 @56", 64, 64, 1, 1, 0, 56, False), ("3x3 64->64 @56", 64, 64, 3, 1, 1, 56, False), ("1x1 64->256 @56 +res", 64, 256, 1, 1, 0, 56, True),
    ("1x1 64->256 @56", 64, 256, 1, 1, 0, 56, False), ("1x1 256->64 @56", 256, 64, 1, 1, 0, 56, False), ("1x1 256->128 @56", 256, 128, 1, 1, 0, 56, False),
    ("3x3s2 128->128 @56", 128, 128, 3, 2, 1, 56, False), ("1x1 128->512 @28 +res", 128, 512, 1, 1, 0, 28, True), ("1x1s2 256->512 @56", 256, 512, 1, 2, 0, 56, False),
    ("1x1 512->128 @28", 512, 128, 1, 1, 0, 28, False), ("3x3 128->128 @28", 128, 128, 3, 1, 1, 28, False), ("1x1 512->256 @28", 512, 256, 1, 1, 0, 28, False),
    ("3x3s2 256->256 @28", 256, 256, 3, 2, 1, 28, False), ("1x1 256->1024 @14 +res", 256, 1024, 1, 1, 0, 14, True), ("1x1s2 512->1024 @28", 512, 1024, 1, 2, 0, 28, False),
    ("1x1 1024->256 @14", 1024, 256, 1, 1, 0, 14, False), ("3x3 256->256 @14", 256, 256, 3, 1, 1, 14, False), ("1x1 1024->512 @14", 1024, 512, 1, 1, 0, 14, False),
    ("3x3s2 512->512 @14", 512, 512, 3, 2, 1, 14, False), ("1x1 512->2048 @7 +res", 512, 2048, 1, 1, 0, 7, True), ("1x1s2 1024->2048 @14", 1024, 2048, 1, 2, 0, 14, False),
    ("1x1 2048->512 @7", 2048, 512, 1, 1, 0, 7, False), ("3x3 512->512 @7", 512, 512, 3, 1, 1, 7, False),
]
COUNT = {"1x1 64->64 @56": 1, "3x3 64->64 @56": 3, "1x1 64->256 @56 +res": 3, "1x1 64->256 @56": 1, "1x1 256->64 @56": 2, "1x1 256->128 @56": 1, "3x3s2 128->128 @56": 1,
         "1x1 128->512 @28 +res": 4, "1x1s2 256->512 @56": 1, "1x1 512->128 @28": 3, "3x3 128->128 @28": 3, "1x1 512->256 @28": 1, "3x3s2 256->256 @28": 1,
         "1x1 256->1024 @14 +res": 6, "1x1s2 512->1024 @28": 1, "1x1 1024->256 @14": 5, "3x3 256->256 @14": 5, "1x1 1024->512 @14": 1, "3x3s2 512->512 @14": 1,
         "1x1 512->2048 @7 +res": 3, "1x1s2 1024->2048 @14": 1, "1x1 2048->512 @7": 2, "3x3 512->512 @7": 2}

which = os.environ.get("PROBE", "layers,kscale")
if "layers" in which:
    tot = 0.0
    for L in LAYERS:
        tot += COUNT[L[0]] * layer_line(*L)
    emit(f"sum over the model's conv layers (stem / fc excluded): {tot:.1f} us")
if "halo" in which:
    os.environ["RTEN_B200_HALO"] = "1"
    os.environ["RTEN_B200_VERBOSE"] = "1"
    for L in LAYERS:
        if L[0].startswith("3x3 "):
            layer_line("halo " + L[0], *L[1:])
    os.environ.pop("RTEN_B200_HALO")
    os.environ.pop("RTEN_B200_VERBOSE")
if "kscale" in which:
    for (co, hw, res) in [(1024, 14, True), (256, 14, False), (256, 56, True), (64, 56, False), (2048, 7, True), (512, 28, True)]:
        emit(f"== K scaling: 1x1 K->{co} @{hw} res={res}")
        for ci in (32, 64, 128, 256, 512, 1024):
            run, fl, by = make(ci, co, 1, 1, 0, hw, res)
            emit(f"   K={ci:5d}: {hot_time(run):6.1f} us")
