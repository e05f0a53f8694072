"""On-box yardsticks (cuBLAS through torch) and a check that CUPTI (torch.profiler) sees the kernels of a replayed
CUDA graph launched by librten_b200.so -- bench.py uses both.  Prints JSON to stdout."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RTEN_B200_F32_MODE", "tf32")

import numpy as np  # noqa: E402
import torch  # noqa: E402

import rten_b200 as rt  # noqa: E402


def measure_matmul_peaks(n=8192, sustained_s=3.0):
    out = {}
    torch.backends.cuda.matmul.allow_tf32 = True
    cases = {
        "bf16": (torch.randn(n, n, device="cuda", dtype=torch.bfloat16), torch.randn(n, n, device="cuda", dtype=torch.bfloat16), torch.matmul),
        "tf32": (torch.randn(n, n, device="cuda"), torch.randn(n, n, device="cuda"), torch.matmul),
        "int8": (torch.randint(-128, 127, (n, n), device="cuda", dtype=torch.int8), torch.randint(-128, 127, (n, n), device="cuda", dtype=torch.int8), torch._int_mm),
    }
    for name, (a, b, fn) in cases.items():
        try:
            for _ in range(3):
                fn(a, b)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(10):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                fn(a, b)
                e.record()
                torch.cuda.synchronize()
                best = min(best, s.elapsed_time(e))
            t0 = time.time()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            cnt = 0
            while time.time() - t0 < sustained_s:
                for _ in range(20):
                    fn(a, b)
                cnt += 20
                torch.cuda.synchronize()
            e.record()
            torch.cuda.synchronize()
            out[name] = {"burst": 2.0 * n ** 3 / best / 1e9, "sustained": 2.0 * n ** 3 * cnt / s.elapsed_time(e) / 1e9}
        except Exception as ex:  # noqa: BLE001
            out[name] = {"error": str(ex)}
    return out


def profiler_sees_graph_kernels():
    from torch.profiler import ProfilerActivity, profile
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = rt.Context(0, stream=stream.cuda_stream)
    a = rt.from_torch(ctx, torch.randn(2048, 768, device="cuda"))
    b = rt.from_torch(ctx, torch.randn(768, 768, device="cuda")).permute(1, 0)
    o = ctx.empty((2048, 768))
    x = rt.from_torch(ctx, torch.randn(2048, 768, device="cuda"))
    g1 = ctx.to_device(np.ones(768, np.float32))
    rt.MatMul().run(ctx, a, b, out=o)
    ctx.sync()
    ctx.graph_begin()
    rt.MatMul().run(ctx, a, b, out=o)
    rt.LayerNormalization(-1, 1e-12).run(ctx, x, g1, g1, out=o)
    g = ctx.graph_end()
    g.launch()
    ctx.sync()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            g.launch()
        torch.cuda.synchronize()
    rows = []
    for ev in prof.key_averages():
        t = getattr(ev, "device_time_total", None)
        if t is None:
            t = getattr(ev, "cuda_time_total", 0.0)
        rows.append({"name": ev.key[:90], "count": ev.count, "device_us_total": t})
    return rows


if __name__ == "__main__":
    res = {"peaks": measure_matmul_peaks()}
    try:
        res["profiler_rows"] = profiler_sees_graph_kernels()
    except Exception as ex:  # noqa: BLE001
        res["profiler_error"] = repr(ex)
    print(json.dumps(res, indent=1))
