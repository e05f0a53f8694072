"""Experiment: ResNet-50 batch 32 as ONE dependency chain (one CUDA graph) vs TWO independent half-batch chains replayed
concurrently on two streams (kernel heads / tails and under-filled grids of one chain overlap the other chain's work).
Single-pass TF32.  Output: gpurun_out/dual_stream.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("RTEN_B200_F32_MODE", "tf32")

import numpy as np  # noqa: E402
import torch  # noqa: E402

import rten_b200 as rt  # noqa: E402
from rten_b200 import graphs  # noqa: E402
from oracle import oracle  # noqa: E402  (synthetic weights only)

torch.cuda.set_device(0)
out = open(os.path.join(ROOT, "gpurun_out", "dual_stream.txt"), "w")


def emit(s):
    print(s, flush=True)
    out.write(s + "\n")
    out.flush()


rng = oracle.XorShiftRng(5678)
spec = graphs.make_resnet50(lambda s: rng.uniform(s))
x = oracle.XorShiftRng(1234).uniform((32, 3, 224, 224))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def build(batch_slices):
    """one context / stream / graph per slice of the batch"""
    parts = []
    for lo, hi in batch_slices:
        stream = torch.cuda.Stream()
        ctx = rt.Context(0, stream=stream.cuda_stream)
        ctx.set_autotune(True)
        runner = graphs.ResNet50Runner(ctx, spec, fuse=True)
        xin = ctx.to_device(x[lo:hi], channels_last=True)
        runner.run(xin)  # warm-up: plans measured
        ctx.sync()
        ctx.set_autotune(False)
        ctx.graph_begin()
        y = runner.run(xin)
        g = ctx.graph_end()
        parts.append((stream, ctx, g, y))
    return parts


def time_parts(parts, steps=20):
    for _ in range(3):
        for s, c, g, y in parts:
            g.launch()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        evs = []
        for s, c, g, y in parts:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s)
            evs.append((a, b))
        for (s, c, g, y), (a, b) in zip(parts, evs):
            g.launch()
            b.record(s)
        torch.cuda.synchronize()
        t0 = evs[0][0]
        ts.append(max(t0.elapsed_time(b) for a, b in evs))
    return float(np.median(ts))


one = build([(0, 32)])
ms1 = time_parts(one)
emit(f"one chain, batch 32:            {ms1 * 1e3:8.1f} us / step = {32 / ms1 * 1e3:8.0f} img/s")
ref = one[0][3].numpy()
for nsplit in (2, 4):
    per = 32 // nsplit
    parts = build([(i * per, (i + 1) * per) for i in range(nsplit)])
    ms = time_parts(parts)
    got = np.concatenate([p[3].numpy() for p in parts], 0)
    err = float(np.abs(got - ref).max() / np.abs(ref).max())
    emit(f"{nsplit} concurrent chains of batch {per:2d}: {ms * 1e3:8.1f} us / step = {32 / ms * 1e3:8.0f} img/s   (logits vs one chain: rel diff {err:.1e})")
