"""ncu target: the 3x3 stride-1 ResNet-50 layers once through the generic implicit-GEMM kernel and once through the
halo-reuse kernel (batch 32, TF32 mode)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RTEN_B200_F32_MODE", "tf32")

import numpy as np  # noqa: E402
import torch  # noqa: E402

import rten_b200 as rt  # noqa: E402

torch.cuda.set_device(0)
ctx = rt.Context(0)
for c, hw in [(64, 56), (128, 28), (256, 14)]:
    x = rt.from_torch(ctx, torch.randn(32, hw, hw, c, device="cuda")).permute(0, 3, 1, 2)
    w = ctx.to_device(np.random.randn(c, c, 3, 3).astype(np.float32))
    bias = ctx.to_device(np.zeros(c, np.float32))
    op = rt.Conv(1, (1, 1), (1, 1, 1, 1), (1, 1), activation=rt.ACT_RELU)
    pk = op.prepack(ctx, 1, w)
    y = op.run(ctx, x, w, bias, packed_w=pk)
    os.environ["RTEN_B200_NO_HALO"] = "1"
    ctx.set_autotune(True)
    op.run(ctx, x, w, bias, packed_w=pk, out=y)
    ctx.set_autotune(False)
    ctx.sync()
    os.environ["NCU_MARK"] = "1"
    op.run(ctx, x, w, bias, packed_w=pk, out=y)   # generic (measured plan)
    os.environ.pop("RTEN_B200_NO_HALO")
    op.run(ctx, x, w, bias, packed_w=pk, out=y)   # halo
    ctx.sync()
print("done")
