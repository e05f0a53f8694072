"""Launch target for `ncu` over the HBM-bound row kernels at the BASELINE shapes (two launches each: the second is the
one to read).  python tools/rowops_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import rten_b200 as rt  # noqa: E402

torch.cuda.set_device(0)
ctx = rt.Context(0)
s4 = rt.from_torch(ctx, torch.randn(16, 12, 128, 128, device="cuda"))
m = ctx.to_device(np.zeros((16, 1, 1, 128), np.float32))
x = rt.from_torch(ctx, torch.randn(2048, 768, device="cuda"))
g = ctx.to_device(np.ones(768, np.float32))
b = ctx.to_device(np.zeros(768, np.float32))
xo = ctx.empty((2048, 768))
h = rt.from_torch(ctx, torch.randn(2048, 3072, device="cuda"))
q8 = rt.from_torch(ctx, torch.randn(4096, 768, device="cuda"))
sg = rt.from_torch(ctx, torch.randn(8, 12, 512, 512, device="cuda"))
mg = ctx.to_device(np.zeros((1, 1, 512, 512), np.float32))
for _ in range(2):
    rt.AddSoftmax().run(ctx, s4, m, in_place=True)
    rt.LayerNormalization(-1, 1e-12).run(ctx, x, g, b, out=xo)
    rt.Gelu().run(ctx, h, in_place=True)
    rt.Gelu(approximate=True).run(ctx, h, in_place=True)
    rt.DynamicQuantizeLinear().run(ctx, q8)
    rt.AddSoftmax().run(ctx, sg, mg, in_place=True)
ctx.sync()
print("done")
