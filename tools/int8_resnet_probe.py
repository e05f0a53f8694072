"""One eager pass of the int8 ResNet-50 runner (batch 64) for `ncu --metrics gpu__time_duration.sum` launch lists."""
import os, sys

os.environ.setdefault("RTEN_B200_F32_MODE", "tf32")  # these tools measure the single-pass TF32 kernels unless told otherwise
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rten_b200 as rt
from rten_b200 import graphs
from oracle import oracle  # weights / inputs RNG only
torch.cuda.set_device(0)
ctx = rt.Context(0)
rng = oracle.XorShiftRng(5678)
q = graphs.quantize_resnet50(graphs.make_resnet50(lambda s: rng.uniform(s)))
x = ctx.to_device(oracle.XorShiftRng(1234).uniform((64, 3, 224, 224)), channels_last=True)
runner = graphs.ResNet50Int8Runner(ctx, q, fuse=True)
for _ in range(int(os.environ.get("PASSES", "2"))):
    y = runner.run(x)
    ctx.sync()
print("done", y.shape)
