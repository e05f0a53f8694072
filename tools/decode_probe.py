"""GPT-2 small int8 b8 (configs[4]): prefill 512, then decode steps replayed from one CUDA graph -- fused decode path
(quantised-linear skinny kernels + single-query attention) vs the separate operators.  Prints tokens/s and launches."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import rten_b200 as rt  # noqa: E402
from oracle import oracle  # noqa: E402  (RNG for the synthetic weights only)
from rten_b200 import graphs  # noqa: E402


def main():
    nsteps = int(os.environ.get("DECODE_STEPS", "32"))
    modes = os.environ.get("DECODE_MODES", "fused,unfused").split(",")
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = rt.Context(0, stream=stream.cuda_stream)
    rng = oracle.XorShiftRng(5678)
    spec = graphs.make_gpt2_int8(lambda s: rng.uniform(s))
    ids = (oracle.XorShiftRng(1).u64(8 * 576) % 50257).astype(np.int32).reshape(8, 576)
    for mode in modes:
        run = graphs.GPT2Int8Runner(ctx, spec, 8, 576)
        ctx.set_autotune(True)
        run.forward(ids[:, :512])
        run.reset()
        ctx.set_autotune(False)
        torch.cuda.synchronize()
        s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        run.forward(ids[:, :512])
        e0.record(stream)
        ctx.set_autotune(True)
        run.build_decode_graph(fused=(mode == "fused"))
        ctx.set_autotune(False)
        torch.cuda.synchronize()
        l0 = ctx.launches
        run.decode_step(ids[:, 512:513])
        per_step = ctx.launches - l0
        # device time of the graph replay alone
        evs = []
        t0 = time.perf_counter()
        for i in range(1, nsteps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            run._write_step_inputs(ids[:, 512 + i:513 + i])
            a.record(stream)
            run._graph.launch()
            b.record(stream)
            run.past += 1
            evs.append((a, b))
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dev_ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
        print(f"{mode}: prefill512 {8 * 512 / (s0.elapsed_time(e0) / 1e3):.0f} tok/s; decode graph replay {dev_ms * 1e3:.1f} us/step "
              f"= {8 / (dev_ms / 1e3):.0f} tok/s device, {8 * (nsteps - 1) / wall:.0f} tok/s incl. host step inputs; {per_step} launches/step",
              flush=True)
        del run


if __name__ == "__main__":
    main()
