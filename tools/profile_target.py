"""One eager pass (no CUDA graph) of a benched model between cudaProfilerStart / cudaProfilerStop, after warm-up passes
that measure (or load) the launch plans -- the target of the ncu captures under profiles/ (run with
`ncu --profile-from-start off ...`).  During the profiled pass RTEN_B200_VERBOSE=1 prints the launch plan of every
tensor-core launch to stderr, in launch order, for tools/layer_table.py.

  python tools/profile_target.py --model resnet50|bert|resnet50_int8|gpt2 [--mode tf32|tf32x3] [--plans FILE]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="resnet50")
ap.add_argument("--mode", default="tf32")
ap.add_argument("--plans", default=None)
args = ap.parse_args()
os.environ["RTEN_B200_F32_MODE"] = args.mode

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import rten_b200 as rt  # noqa: E402
from rten_b200 import graphs  # noqa: E402
from oracle import oracle  # noqa: E402  (synthetic weights / inputs only)

torch.cuda.set_device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = rt.Context(0, stream=stream.cuda_stream)
if args.plans and os.path.exists(args.plans):
    ctx.load_plans(args.plans)
ctx.set_autotune(True)
model = args.model
batch = bench.MODELS[model]["batch"]
spec = bench.make_spec(oracle, model)
inp = bench.make_inputs(oracle, model, batch)
if model == "resnet50":
    runner = graphs.ResNet50Runner(ctx, spec, fuse=True)
    x = ctx.to_device(inp["x"], channels_last=True)
    step = lambda: runner.run(x)
elif model == "resnet50_int8":
    runner = graphs.ResNet50Int8Runner(ctx, spec, fuse=True)
    x = ctx.to_device(inp["x"], channels_last=True)
    step = lambda: runner.run(x)
elif model == "bert":
    runner = graphs.BertRunner(ctx, spec, fuse=True)
    ids, tt, mask = ctx.to_device(inp["ids"]), ctx.to_device(inp["tt"]), ctx.to_device(inp["mask"])
    step = lambda: runner.run(ids, tt, mask)
else:
    runner = graphs.GPT2Int8Runner(ctx, spec, batch, bench.GPT2_CACHE)
    ids = inp["ids"]
    runner.forward(ids[:, :bench.GPT2_PREFILL])  # 512-token prefill fills the KV cache
    runner.build_decode_graph(fused=True)
    tok = ids[:, bench.GPT2_PREFILL:bench.GPT2_PREFILL + 1]

    def step():  # one decode step, launches issued one by one (the same launch list the graph replays)
        runner._write_step_inputs(tok)
        runner._decode_fixed()
for _ in range(3):
    step()
ctx.sync()
ctx.set_autotune(False)
if args.plans:
    ctx.save_plans(args.plans)
step()
ctx.sync()
os.environ["RTEN_B200_VERBOSE"] = "1"
print("== profiled pass ==", file=sys.stderr, flush=True)
torch.cuda.cudart().cudaProfilerStart()
step()
ctx.sync()
torch.cuda.cudart().cudaProfilerStop()
os.environ.pop("RTEN_B200_VERBOSE")
print("done", flush=True)
