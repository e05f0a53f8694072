"""Per-launch report of one ResNet-50 (or BERT) pass: tile configuration chosen by the host (RTEN_B200_VERBOSE) next to
the CUDA-event time of each tensor-core op (graph-less, so tiny ops include launch gaps; use for relative ranking)."""
import os

os.environ.setdefault("RTEN_B200_F32_MODE", "tf32")  # these tools measure the single-pass TF32 kernels unless told otherwise
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

if "--child" not in sys.argv:
    env = dict(os.environ, RTEN_B200_VERBOSE="1")
    p = subprocess.run([sys.executable, __file__, "--child"] + sys.argv[1:], env=env, capture_output=True, text=True)
    cfg = [l for l in p.stderr.splitlines() if l.startswith("[umma_gemm]")]
    times = [l for l in p.stdout.splitlines() if l.startswith("T ")]
    n = len(times)
    cfg = cfg[-n:]  # configurations of the last (timed) pass
    tot = 0.0
    for c, t in zip(cfg, times):
        _, ms, fl = t.split()
        tot += float(ms)
        print(f"{float(ms)*1e3:8.1f} us {float(fl)/float(ms)/1e9:7.1f} TF/s  {c[12:]}")
    print(f"sum of tensor-core ops: {tot*1e3:.1f} us")
    print(p.stdout.splitlines()[-1] if p.stdout else p.stderr[-2000:])
    sys.exit(0)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import rten_b200 as rt  # noqa: E402
import rten_b200.ops as O  # noqa: E402
from oracle import oracle  # noqa: E402
from rten_b200 import graphs  # noqa: E402

model = "bert" if "bert" in sys.argv else "resnet50"
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = rt.Context(0, stream=stream.cuda_stream)
rng = oracle.XorShiftRng(5678)
if model == "resnet50":
    spec = graphs.make_resnet50(lambda s: rng.uniform(s))
    runner = graphs.ResNet50Runner(ctx, spec)
    x = ctx.to_device(oracle.XorShiftRng(1234).uniform((32, 3, 224, 224)), channels_last=True)
    step = lambda: runner.run(x)
else:
    spec = graphs.make_bert(lambda s: rng.uniform(s))
    runner = graphs.BertRunner(ctx, spec)
    ids = ctx.to_device((oracle.XorShiftRng(1234).u64(16 * 128) % 30522).astype(np.int32).reshape(16, 128))
    tt = ctx.to_device(np.zeros((16, 128), np.int32))
    mask = ctx.to_device(np.zeros((16, 1, 1, 128), np.float32))
    step = lambda: runner.run(ids, tt, mask)
records = []


def wrap(orig, flops_fn):
    def run(self, c, *a, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record(stream)
        y = orig(self, c, *a, **k)
        e.record(stream)
        records.append((s, e, flops_fn(a, y)))
        return y
    return run


def conv_flops(a, y):
    w = a[1]
    b, o, oh, ow = y.shape
    return 2.0 * b * o * oh * ow * w.shape[1] * w.shape[2] * w.shape[3]


def mm_flops(a, y):
    return 2.0 * float(np.prod(y.shape)) * a[0].shape[-1]


O.Conv.run = wrap(O.Conv.run, conv_flops)
O.FusedMatMul.run = wrap(O.FusedMatMul.run, mm_flops)
O.MatMul.run = wrap(O.MatMul.run, mm_flops)
O.Gemm.run = wrap(O.Gemm.run, lambda a, y: 2.0 * float(np.prod(y.shape)) * a[0].shape[-1])
for rep in range(3):
    records.clear()
    step()
    torch.cuda.synchronize()
for s, e, f in records:
    print("T", s.elapsed_time(e), f)
print("ops", len(records))
