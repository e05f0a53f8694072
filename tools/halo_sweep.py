"""3x3 stride-1 ResNet-50 layers (batch 32): generic implicit-GEMM kernel (autotuned) vs the halo-reuse kernel over its
unit shapes (RTEN_B200_HALO_BN / _T), 10 chained launches per CUDA-graph replay (hot operands).  Output: gpurun_out/halo_sweep.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RTEN_B200_F32_MODE", "tf32")
os.environ["RTEN_B200_HALO"] = "1"

import numpy as np  # noqa: E402
import torch  # noqa: E402

import rten_b200 as rt  # noqa: E402


def main():
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = rt.Context(0, stream=stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = open(os.path.join(ROOT, "gpurun_out", "halo_sweep.txt"), "w")

    def emit(s):
        print(s, flush=True)
        out.write(s + "\n")
        out.flush()

    def timeit(fn, reps=10):
        # `reps` chained launches per graph replay (programmatic dependent launch, as in the model graph): hot operands,
        # sub-microsecond resolution per launch
        fn()
        ctx.graph_begin()
        for _ in range(reps):
            fn()
        g = ctx.graph_end()
        for _ in range(2):
            g.launch()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            g.launch()
            b.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / reps)
        return float(np.median(ts)) * 1e3

    B = int(os.environ.get("HALO_BATCH", "32"))
    for name, c, hw in [("3x3 64->64 @56", 64, 56), ("3x3 128->128 @28", 128, 28), ("3x3 256->256 @14", 256, 14), ("3x3 512->512 @7", 512, 7)]:
        x = rt.from_torch(ctx, torch.randn(B, hw, hw, c, device="cuda")).permute(0, 3, 1, 2)
        w = ctx.to_device(np.random.randn(c, c, 3, 3).astype(np.float32))
        bias = ctx.to_device(np.zeros(c, np.float32))
        op = rt.Conv(1, (1, 1), (1, 1, 1, 1), (1, 1), activation=rt.ACT_RELU)
        pk = op.prepack(ctx, 1, w)
        y = op.run(ctx, x, w, bias, packed_w=pk)
        fl = 2.0 * B * c * hw * hw * c * 9
        run = lambda: op.run(ctx, x, w, bias, packed_w=pk, out=y)
        os.environ["RTEN_B200_NO_HALO"] = "1"
        ctx.set_autotune(True)
        run()
        ctx.set_autotune(False)
        us = timeit(run)
        os.environ.pop("RTEN_B200_NO_HALO")
        emit(f"== {name} (batch {B}, {fl / 1e9:.2f} GFLOP): generic autotuned {us:.1f} us = {fl / us / 1e6:.0f} TF/s")
        us = timeit(run)
        emit(f"   halo (model's choice)  {us:7.1f} us = {fl / us / 1e6:5.0f} TF/s")
        for bn in (32, 64, 128, 256):
            if bn > c:
                continue
            for T in (1, 2, 3, 4):
                if T * bn > 512:
                    continue
                os.environ["RTEN_B200_HALO_BN"], os.environ["RTEN_B200_HALO_T"] = str(bn), str(T)
                try:
                    us = timeit(run)
                    emit(f"   halo bn={bn:3d} T={T}  {us:7.1f} us = {fl / us / 1e6:5.0f} TF/s")
                except rt.OpError as e:
                    emit(f"   halo bn={bn} T={T}: {e}")
        os.environ.pop("RTEN_B200_HALO_BN", None)
        os.environ.pop("RTEN_B200_HALO_T", None)


if __name__ == "__main__":
    main()
