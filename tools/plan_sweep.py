"""Times forced launch plans (RTEN_B200_FORCE_*) of the tcgen05 conv kernel on the ResNet-50 layers that under-fill the
148 SMs, replayed from CUDA graphs with the L2 flushed between launches.  Output: gpurun_out/plan_sweep.txt"""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RTEN_B200_F32_MODE", "tf32")

import numpy as np  # noqa: E402
import torch  # noqa: E402

import rten_b200 as rt  # noqa: E402

KEYS = ("BN", "PAIR", "KATOMS", "SPLITK", "CTA2")


def main():
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = rt.Context(0, stream=stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    B = 32
    layers = [("3x3 256->256 @14", 256, 256, 3, 1, 1, 14), ("3x3 512->512 @7", 512, 512, 3, 1, 1, 7), ("1x1 256->1024 @14", 256, 1024, 1, 1, 0, 14),
              ("1x1 1024->256 @14", 1024, 256, 1, 1, 0, 14), ("1x1 2048->512 @7", 2048, 512, 1, 1, 0, 7), ("1x1 512->2048 @7", 512, 2048, 1, 1, 0, 7),
              ("3x3 128->128 @28", 128, 128, 3, 1, 1, 28), ("3x3 64->64 @56", 64, 64, 3, 1, 1, 56)]
    out = open(os.path.join(ROOT, "gpurun_out", "plan_sweep.txt"), "w")

    def emit(s):
        print(s, flush=True)
        out.write(s + "\n")
        out.flush()

    for name, ci, co, k, s, p, hw in layers:
        x = rt.from_torch(ctx, torch.randn(B, hw, hw, ci, device="cuda")).permute(0, 3, 1, 2)
        w = ctx.to_device(np.random.randn(co, ci, k, k).astype(np.float32))
        bias = ctx.to_device(np.zeros(co, np.float32))
        op = rt.Conv(1, (1, 1), (p, p, p, p), (s, s), activation=rt.ACT_RELU)
        pk = op.prepack(ctx, 1, w)
        y = op.run(ctx, x, w, bias, packed_w=pk)
        fl = 2.0 * B * co * y.shape[2] * y.shape[3] * ci * k * k
        emit(f"== {name}: {fl / 1e9:.2f} GFLOP")
        results = []
        for bn, pair, katoms, sk, cta2 in itertools.product((64, 128, 192, 256), (0, 1), (1, 2), (1, 2, 3, 4, 6, 8), (0, 1)):
            if bn > co:
                continue
            for kname, v in zip(KEYS, (bn, pair, katoms, sk, cta2)):
                os.environ["RTEN_B200_FORCE_" + kname] = str(v)
            h0, m0 = ctx.forced_plan_counts()
            try:
                op.run(ctx, x, w, bias, packed_w=pk, out=y)
                h1, m1 = ctx.forced_plan_counts()
                if h1 == h0:
                    continue  # no valid plan with these fields
                ctx.graph_begin()
                op.run(ctx, x, w, bias, packed_w=pk, out=y)
                g = ctx.graph_end()
            except rt.OpError as e:
                emit(f"   bn={bn} pair={pair} katoms={katoms} splitk={sk} cta2={cta2}: {e}")
                continue
            for _ in range(2):
                g.launch()
            ts = []
            for _ in range(7):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                g.launch()
                b.record(stream)
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            us = float(np.median(ts)) * 1e3
            results.append((us, bn, pair, katoms, sk, cta2))
            del g
        results.sort()
        for us, bn, pair, katoms, sk, cta2 in results[:12]:
            emit(f"   {us:8.1f} us {fl / us / 1e6:7.1f} TF/s  bn={bn} pair={pair} katoms={katoms} splitk={sk} cta2={cta2}")
        best_nosplit = min((r for r in results if r[4] == 1), default=None)
        if best_nosplit:
            emit(f"   best without split-K: {best_nosplit[0]:.1f} us  bn={best_nosplit[1]} pair={best_nosplit[2]} katoms={best_nosplit[3]} cta2={best_nosplit[5]}")
    for kname in KEYS:
        os.environ.pop("RTEN_B200_FORCE_" + kname, None)


if __name__ == "__main__":
    main()
