#!/usr/bin/env python
"""bench.py -- one "step" = one pass of the hot path (the post-fusion ResNet-50 fp32 op list, batch 32 per
GPU: BASELINE.json configs[1]) over one batch of synthetic input.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model resnet50|bert]

Prints ONE JSON line (rank 0).  `value` = whole-job images/s with inputs resident in HBM; `e2e` = the same
metric through the public operator API with HOST (pinned) input and output buffers, host<->device copies
inside the timed region.  `roofline` describes the dominant kernel (the tcgen05 implicit-GEMM conv),
`cpu_baseline` the CPU restatement of the reference path (oracle/) on a bounded sample.
`--impl reference` times that CPU restatement alone (the Rust reference cannot be built here: no cargo).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons sampled every ~5 ms DURING the timed region through NVML (falls back to
    `nvidia-smi -lms` when the binding is missing)."""

    def __init__(self, index: int):
        self.index, self.rows, self.stop_flag, self.thread = index, [], False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, mx, rs))
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.nv is None:
            return
        self.stop_flag = False
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "note": "NVML binding unavailable: clocks not sampled"}
        self.stop_flag = True
        self.thread.join(timeout=1)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        reasons = set()
        for _, _, rs in self.rows:
            for n, bit in names.items():
                if rs & bit:
                    reasons.add(n)
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(oracle, model, batch):
    rng = oracle.XorShiftRng(1234)
    if model == "resnet50":
        return {"x": rng.uniform((batch, 3, 224, 224))}
    ids = (rng.u64(batch * 128) % 30522).astype(np.int32).reshape(batch, 128)
    return {"ids": ids, "tt": np.zeros((batch, 128), np.int32), "mask": np.zeros((batch, 1, 1, 128), np.float32)}


def make_spec(oracle, model):
    from rten_b200 import graphs
    rng = oracle.XorShiftRng(5678)
    if model == "resnet50":
        return graphs.make_resnet50(lambda s: rng.uniform(s))
    return graphs.make_bert(lambda s: rng.uniform(s))


def run_reference_arm(args, model, batch):
    """CPU restatement of the reference path on all host threads, bounded sample per step."""
    from oracle import oracle
    import model_ref
    ncores = oracle.use_all_cores()
    spec = make_spec(oracle, model)
    # images are the outer parallel level: a many-core host needs the whole batch in flight to be busy
    sample = (batch if ncores >= 16 else 8) if model == "resnet50" else (batch if ncores >= 16 else 4)
    inp = make_inputs(oracle, model, sample)
    arena = oracle.Arena()  # = the reference's BufferPool: operator outputs are recycled from pass to pass
    run = (lambda: model_ref.resnet50_oracle(oracle, spec, inp["x"], arena)) if model == "resnet50" else \
        (lambda: model_ref.bert_oracle(oracle, spec, inp["ids"], inp["tt"], inp["mask"]))
    for _ in range(max(1, min(args.warmup, 1))):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    unit = "img/s" if model == "resnet50" else "seq/s"
    cores = oracle.num_threads()
    return {
        "impl": "reference", "metric": metric_name(model), "value": val, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_of(model, batch, args.gpus),
        "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": "port",
                         "sample": f"{sample} of {batch} inputs per step; CPU restatement of the rten path (oracle/), the Rust reference cannot be built here"},
        "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def metric_name(model):
    return "resnet50_fp32_inferences_per_sec" if model == "resnet50" else "bert_base_fp32_seq128_inferences_per_sec"


def config_of(model, batch, n):
    if model == "resnet50":
        return {"workload": "ResNet-50 fp32 (post-fusion op list, BN folded), batch 32 per GPU, 224x224, synthetic weights XorShift(5678)",
                "global_batch": batch * n, "per_gpu_batch": batch, "parallelism": f"dp{n} (batch shard, all-gather of logits)",
                "f32_mode": "tf32 single pass", "l2": "256 MiB memset between timed steps"}
    return {"workload": "BERT-base fp32 (post-fusion op list), batch 16 x seq 128 per GPU, synthetic weights XorShift(5678)",
            "global_batch": batch * n, "per_gpu_batch": batch, "seq_len": 128, "parallelism": f"dp{n} (batch shard, all-gather of hidden states)",
            "f32_mode": "tf32 single pass", "l2": "256 MiB memset between timed steps"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="resnet50", choices=["resnet50", "bert"])
    ap.add_argument("--no-graph", action="store_true", help="issue ops one by one instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-autotune", action="store_true", help="use the cost model's launch plans instead of timing candidates during warm-up")
    ap.add_argument("--plans", default=None, help="file of measured launch plans: loaded if it exists, (re)written after the warm-up pass")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary numbers (BERT-base pass, 8192^3 GEMM TFLOP/s)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    model = args.model
    batch = 32 if model == "resnet50" else 16
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(run_reference_arm(args, model, batch)), flush=True)
        return

    import torch
    import torch.distributed as dist
    import rten_b200 as rt
    from rten_b200 import graphs, shard
    from oracle import oracle  # inputs/weights RNG + cpu_baseline leg only

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: rten_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = rt.Context(local_rank, stream=stream.cuda_stream)
    ctx.set_autotune(not args.no_autotune)  # plans are measured during the first (untimed, eager) pass
    if args.plans and os.path.exists(args.plans):
        ctx.load_plans(args.plans)

    spec = make_spec(oracle, model)
    inp = make_inputs(oracle, model, batch)
    if model == "resnet50":
        runner = graphs.ResNet50Runner(ctx, spec, fuse=True)
        x_dev = ctx.to_device(inp["x"], channels_last=True)
        dev_inputs = [x_dev]
        step_fn = lambda: runner.run(x_dev)
        flops = graphs.resnet50_flops(spec) * batch
        unit = "img/s"
    else:
        runner = graphs.BertRunner(ctx, spec, fuse=True)
        ids, tt, mask = ctx.to_device(inp["ids"]), ctx.to_device(inp["tt"]), ctx.to_device(inp["mask"])
        dev_inputs = [ids, tt, mask]
        step_fn = lambda: runner.run(ids, tt, mask)
        flops = graphs.bert_flops(spec, batch, 128)
        unit = "seq/s"

    # ---- eager run (also warms the buffer pool so that graph capture never allocates)
    out = step_fn()
    ctx.sync()
    if args.plans and rank == 0:
        ctx.save_plans(args.plans)
    out_shape = out.shape
    del out
    gather_buf = torch.empty(shard.gather_layout(world, tuple(out_shape)), dtype=torch.float32, device="cuda") if world > 1 else None
    out_t = torch.empty(tuple(out_shape), dtype=torch.float32, device="cuda")
    out_dst = rt.from_torch(ctx, out_t)

    def copy_out(o):
        import ctypes as C
        src, dst = o.desc(), out_dst.desc()
        ctx.check(ctx.lib.rten_b200_copy(ctx.handle, C.byref(src), C.byref(dst)))

    launches_per_step = None
    graph = None
    if not args.no_graph:
        l0 = ctx.launches
        ctx.graph_begin()
        o = step_fn()
        copy_out(o)
        graph = ctx.graph_end()
        del o

    def device_step():
        if graph is not None:
            graph.launch()
        else:
            copy_out(step_fn())
        if world > 1:
            shard.all_gather_outputs(dist, out_t, gather_buf)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        evs = []
        l0 = ctx.launches
        for _ in range(steps):
            flush.zero_()  # L2 flush, outside the timed events
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream)
            fn()
            e.record(stream)
            evs.append((s, e))
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        if world > 1:
            dist.barrier()
        ms = sum(s.elapsed_time(e) for s, e in evs)
        launches = ctx.launches - l0
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, clocks

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, launches, clocks = timed(device_step, args.steps, args.warmup, sampler)
    value = batch * world * args.steps / (ms / 1e3)

    # ---- e2e: host (pinned) inputs -> H2D -> op list -> D2H of the result, all inside the timed region
    import ctypes as C
    pinned = []
    order = ["x"] if model == "resnet50" else ["ids", "tt", "mask"]
    for name, d in zip(order, dev_inputs):
        h = ctx.pinned_empty(inp[name].shape, inp[name].dtype)
        h[...] = inp[name]
        pinned.append((h, d))
    host_out = ctx.pinned_empty(out_shape, np.float32)
    h2d = sum(h.nbytes for h, _ in pinned)
    d2h = host_out.nbytes

    # Double-buffered pipeline through the public API: a second context bound to a copy stream moves step i+1's input
    # from pinned host memory into a raw device buffer while step i computes; the compute stream re-lays it out into the
    # model's input tensor (channels-last for ResNet-50), replays the step and copies the result back to the host.
    # EVERY step's H2D and D2H are inside the timed region; nothing is reused across steps.
    copy_stream = torch.cuda.Stream()
    cctx = rt.Context(local_rank, stream=copy_stream.cuda_stream)
    raw = [[cctx.empty(h.shape, h.dtype) for h, _ in pinned] for _ in range(2)]
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_consumed = [torch.cuda.Event() for _ in range(2)]
    host_outs = [host_out, ctx.pinned_empty(out_shape, np.float32)]

    def issue_h2d(i):
        b = i % 2
        if i >= 2:
            copy_stream.wait_event(ev_consumed[b])
        for (h, _), r in zip(pinned, raw[b]):
            src = rt.ops._desc(h.ctypes.data, h.dtype, h.shape, rt.ops._contig(h.shape), -1)
            dst = r.desc()
            cctx.check(cctx.lib.rten_b200_copy(cctx.handle, C.byref(src), C.byref(dst)))
        ev_copied[b].record(copy_stream)

    out_bufs = [ctx.empty(out_shape, np.float32) for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]

    def issue_d2h(i):
        # result of step i: device copy made by the compute stream -> pinned host buffer, on the copy stream
        b = i % 2
        copy_stream.wait_event(ev_done[b])
        src = out_bufs[b].desc()
        ho = host_outs[b]
        dst = rt.ops._desc(ho.ctypes.data, ho.dtype, ho.shape, rt.ops._contig(ho.shape), -1)
        cctx.check(cctx.lib.rten_b200_copy(cctx.handle, C.byref(src), C.byref(dst)))

    def e2e_run(steps):
        """-> device milliseconds from the first H2D to the last D2H of `steps` pipelined steps"""
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(copy_stream)
        issue_h2d(0)
        for i in range(steps):
            b = i % 2
            stream.wait_event(ev_copied[b])
            for (_, d), r in zip(pinned, raw[b]):
                a_, b_ = r.desc(), d.desc()
                ctx.check(ctx.lib.rten_b200_copy(ctx.handle, C.byref(a_), C.byref(b_)))  # layout change, device to device
            ev_consumed[b].record(stream)
            flush.zero_()  # L2 flush between steps (inside the timed region here)
            device_step()
            a_, b_ = out_dst.desc(), out_bufs[b].desc()
            ctx.check(ctx.lib.rten_b200_copy(ctx.handle, C.byref(a_), C.byref(b_)))
            ev_done[b].record(stream)
            # the host now feeds the NEXT step and collects the PREVIOUS result while this step runs
            if i + 1 < steps:
                issue_h2d(i + 1)
            if i >= 1:
                issue_d2h(i - 1)
        issue_d2h(steps - 1)
        t1.record(copy_stream)
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    e2e_run(3)
    ms_e2e = e2e_run(args.steps)
    e2e_value = batch * world * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel: per-launch CUDA-event timing of every GEMM/conv op of one pass
    roof = None
    if rank == 0:
        roof = roofline_pass(ctx, rt, runner, model, spec, dev_inputs, batch, stream, torch)

    extras = None
    if rank == 0 and not args.no_extras:
        extras = secondary_numbers(ctx, rt, graphs, oracle, model, stream, torch, flush)

    if rank == 0:
        peaks = load_peaks()
        line = {
            "metric": metric_name(model), "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32(tf32 mma)",
            "data": "synthetic", "config": config_of(model, batch, world), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / args.steps,
                    "how": "double-buffered: the copy stream moves step i+1's input H2D and step i-1's result D2H while step i computes; every step's H2D + D2H and the L2 flush are inside the timed region"},
            "gpu_launches": int(launches),
            "cuda_graph": graph is not None,
            "model_tflops": flops * world * args.steps / (ms / 1e3) / 1e12,
        }
        if roof:
            tf32_peak = 0.5 * peaks["bf16_sustained"]
            roof_line = {"bound": "tensor", "kernel": "rtb::umma_gemm_kernel<0> (tcgen05 kind::tf32 implicit-GEMM conv / GEMM)",
                         "achieved": roof["tflops"], "peak": tf32_peak, "unit": "TFLOP/s", "frac": roof["tflops"] / tf32_peak,
                         "traffic": ncu_traffic(), "launches_timed": roof["launches"], "share_of_step": roof["share"],
                         "peak_source": f"0.5 x {peaks['src']} bf16 sustained ({peaks['bf16_sustained']} TF/s): kind::tf32 issues at half the bf16 rate",
                         # the same launches against the other roof: at batch 32 the wide 1x1 layers are nearer to it
                         "hbm_view": {"achieved": roof["gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                      "frac": roof["gbs"] / peaks["hbm_gbs"], "bytes": "algorithmic: operands + output (+ residual) once"}}
            line["roofline"] = roof_line
        if extras:
            tf32_peak = 0.5 * peaks["bf16_burst"]
            extras["gemm_tf32_8192_frac_of_peak"] = extras["gemm_tf32_8192_tflops"] / tf32_peak
            extras["gemm_int8_8192_frac_of_peak"] = extras["gemm_int8_8192_tops"] / (2.0 * peaks["bf16_burst"])
            extras["peaks"] = f"tf32 = 0.5 x, int8 = 2 x {peaks['src']} bf16 burst ({peaks['bf16_burst']} TF/s): kernels timed alone"
            line["also"] = extras
        if not args.no_cpu_baseline:
            a2 = argparse.Namespace(**vars(args))
            a2.steps, a2.warmup = 1, 1
            ref = run_reference_arm(a2, model, batch)
            line["cpu_baseline"] = ref["cpu_baseline"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu capture of
    this same command (profiles/r01_ncu_resnet50.json); None if that summary is absent."""
    p = os.path.join(ROOT, "profiles", "r01_ncu_resnet50.json")
    try:
        return json.load(open(p))["umma_avg_dram_bytes_per_launch"]
    except Exception:
        return None


def secondary_numbers(ctx, rt, graphs, oracle, model, stream, torch, flush):
    """Secondary numbers the BASELINE metric names (BERT-base pass, GEMM TFLOP/s); same timing hygiene, few steps."""
    out = {}

    def timed(fn, iters=5, warm=2):
        fn()
        ctx.graph_begin()
        fn()
        g = ctx.graph_end()
        for _ in range(warm):
            g.launch()
        ms = []
        for _ in range(iters):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream)
            g.launch()
            e.record(stream)
            torch.cuda.synchronize()
            ms.append(s.elapsed_time(e))
        return float(np.median(ms))

    n = 8192
    a = rt.from_torch(ctx, torch.randn(n, n, device="cuda"))
    b = rt.from_torch(ctx, torch.randn(n, n, device="cuda")).permute(1, 0)
    o = ctx.empty((n, n))
    ms = timed(lambda: rt.MatMul().run(ctx, a, b, out=o))
    out["gemm_tf32_8192_tflops"] = 2.0 * n ** 3 / ms / 1e9
    ai = rt.from_torch(ctx, torch.randint(0, 255, (n, n), device="cuda", dtype=torch.uint8))
    bi = rt.from_torch(ctx, torch.randint(-128, 127, (n, n), device="cuda", dtype=torch.int8)).permute(1, 0)
    oi = ctx.empty((n, n), np.int32)
    ms = timed(lambda: rt.MatMulInteger().run(ctx, ai, bi, out=oi))
    out["gemm_int8_8192_tops"] = 2.0 * n ** 3 / ms / 1e9
    del a, b, o, ai, bi, oi
    other = "bert" if model == "resnet50" else "resnet50"
    spec = make_spec(oracle, other)
    inp = make_inputs(oracle, other, 16 if other == "bert" else 32)
    if other == "bert":
        runner = graphs.BertRunner(ctx, spec)
        ids, tt, mask = ctx.to_device(inp["ids"]), ctx.to_device(inp["tt"]), ctx.to_device(inp["mask"])
        ms = timed(lambda: runner.run(ids, tt, mask))
        out["bert_base_fp32_b16_s128_seq_per_sec"] = 16 / (ms / 1e3)
        out["bert_base_model_tflops"] = graphs.bert_flops(spec, 16, 128) / ms / 1e9
    else:
        runner = graphs.ResNet50Runner(ctx, spec)
        x = ctx.to_device(inp["x"], channels_last=True)
        ms = timed(lambda: runner.run(x))
        out["resnet50_fp32_b32_img_per_sec"] = 32 / (ms / 1e3)
    del runner
    # configs[3]: dynamically quantised ResNet-50, batch 64 (DynamicQuantizeLinear -> ConvIntegerToFloat(+bias, +identity, Relu))
    rspec = spec if other == "resnet50" else make_spec(oracle, "resnet50")
    qrunner = graphs.ResNet50Int8Runner(ctx, graphs.quantize_resnet50(rspec), fuse=True)
    x64 = ctx.to_device(make_inputs(oracle, "resnet50", 64)["x"], channels_last=True)
    ms = timed(lambda: qrunner.run(x64))
    out["resnet50_int8_b64_img_per_sec"] = 64 / (ms / 1e3)
    del qrunner, x64
    # configs[4]: GPT-2 small int8, batch 8: prefill of 512 tokens, then decode steps against the KV cache (eager
    # launches: the cache length changes every step)
    grng = oracle.XorShiftRng(5678)
    gspec = graphs.make_gpt2_int8(lambda s: grng.uniform(s))
    grun = graphs.GPT2Int8Runner(ctx, gspec, 8, 576)
    gids = (oracle.XorShiftRng(1).u64(8 * 576) % 50257).astype(np.int32).reshape(8, 576)
    grun.forward(gids[:, :512])  # warm-up (autotune, pool)
    grun.forward(gids[:, 512:513])
    grun.reset()
    ctx.set_autotune(False)  # keep the measured plans, stop measuring: the attention shapes change every decode step
    torch.cuda.synchronize()
    s0, e0 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
    s0.record(stream)
    grun.forward(gids[:, :512])
    e0.record(stream)
    ndec = 32
    ctx.set_autotune(True)
    grun.build_decode_graph()  # untimed: capture of the fixed-shape decode step
    ctx.set_autotune(False)
    torch.cuda.synchronize()
    e0b, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0b.record(stream)
    for i in range(ndec):
        grun.decode_step(gids[:, 512 + i:513 + i])  # per step: two small H2D copies + one graph replay
    e1.record(stream)
    torch.cuda.synchronize()
    out["gpt2_int8_b8_prefill512_tokens_per_sec"] = 8 * 512 / (s0.elapsed_time(e0) / 1e3)
    out["gpt2_int8_b8_decode_tokens_per_sec"] = 8 * ndec / (e0b.elapsed_time(e1) / 1e3)
    return out


def roofline_pass(ctx, rt, runner, model, spec, dev_inputs, batch, stream, torch):
    """Time each tensor-core op of one eager pass with CUDA events on the launching stream (3 repetitions, after
    warm-up) and divide the algorithmic flops by the summed durations."""
    import rten_b200.ops as O
    records = []
    orig_conv, orig_mm, orig_mm0 = O.Conv.run, O.FusedMatMul.run, O.MatMul.run

    def nbytes(t):
        return float(np.prod(t.shape)) * 4.0 if t is not None and hasattr(t, "shape") else 0.0

    def wrap(orig, flops_fn):
        def run(self, c, *a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream)
            y = orig(self, c, *a, **k)
            e.record(stream)
            # algorithmic bytes: every operand read once, the output written once
            by = nbytes(a[0]) + nbytes(a[1]) + nbytes(y) + nbytes(k.get("residual")) + (nbytes(a[2]) if len(a) > 2 else 0.0)
            records.append((s, e, flops_fn(a, y), by))
            return y
        return run

    def conv_flops(a, y):
        w = a[1]
        b, o, oh, ow = y.shape
        return 2.0 * b * o * oh * ow * w.shape[1] * w.shape[2] * w.shape[3]

    def mm_flops(a, y):
        k = a[0].shape[-1]
        return 2.0 * float(np.prod(y.shape)) * k

    O.Conv.run = wrap(orig_conv, conv_flops)
    O.FusedMatMul.run = wrap(orig_mm, mm_flops)
    O.MatMul.run = wrap(orig_mm0, mm_flops)
    try:
        tot_ms, tot_fl, tot_by, n = 0.0, 0.0, 0.0, 0
        for rep in range(4):
            records.clear()
            s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(stream)
            if model == "resnet50":
                runner.run(dev_inputs[0])
            else:
                runner.run(*dev_inputs)
            e0.record(stream)
            torch.cuda.synchronize()
            if rep == 0:
                continue
            tot_ms += sum(s.elapsed_time(e) for s, e, _, _ in records)
            tot_fl += sum(f for _, _, f, _ in records)
            tot_by += sum(b for _, _, _, b in records)
            n += len(records)
            step_ms = s0.elapsed_time(e0)
            share = sum(s.elapsed_time(e) for s, e, _, _ in records) / step_ms
    finally:
        O.Conv.run, O.FusedMatMul.run, O.MatMul.run = orig_conv, orig_mm, orig_mm0
    return {"tflops": tot_fl / (tot_ms / 1e3) / 1e12, "gbs": tot_by / (tot_ms / 1e3) / 1e9, "launches": n, "share": share}


if __name__ == "__main__":
    main()
