"""A few representative tensor-core launches for `ncu --set full` (no CUDA graph, 2 launches each)."""
import os

os.environ.setdefault("RTEN_B200_F32_MODE", "tf32")  # these tools measure the single-pass TF32 kernels unless told otherwise
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import rten_b200 as rt  # noqa: E402

torch.cuda.set_device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = rt.Context(0, stream=stream.cuda_stream)
which = sys.argv[1:] or ["c64_256", "c256_64", "c3x3_64", "gemm4096", "igemm4096"]
B = 32
for w in which:
    if w.startswith("c"):
        ci, co, k, s, p, hw = {"c64_256": (64, 256, 1, 1, 0, 56), "c256_64": (256, 64, 1, 1, 0, 56), "c3x3_64": (64, 64, 3, 1, 1, 56),
                               "c3x3_512": (512, 512, 3, 1, 1, 7), "c512_2048": (512, 2048, 1, 1, 0, 7)}[w]
        x = rt.from_torch(ctx, torch.randn(B, hw, hw, ci, device="cuda")).permute(0, 3, 1, 2)
        wt = ctx.to_device(np.random.randn(co, ci, k, k).astype(np.float32))
        bias = ctx.to_device(np.zeros(co, np.float32))
        op = rt.Conv(1, (1, 1), (p, p, p, p), (s, s), activation=rt.ACT_RELU)
        pk = op.prepack(ctx, 1, wt)
        y = op.run(ctx, x, wt, bias, packed_w=pk)
        for _ in range(2):
            op.run(ctx, x, wt, bias, packed_w=pk, out=y)
    elif w == "gemm4096":
        a = rt.from_torch(ctx, torch.randn(4096, 4096, device="cuda"))
        b = rt.from_torch(ctx, torch.randn(4096, 4096, device="cuda")).permute(1, 0)
        out = ctx.empty((4096, 4096))
        for _ in range(3):
            rt.MatMul().run(ctx, a, b, out=out)
    elif w == "igemm4096":
        a = rt.from_torch(ctx, torch.randint(0, 255, (4096, 4096), device="cuda", dtype=torch.uint8))
        b = rt.from_torch(ctx, torch.randint(-128, 127, (4096, 4096), device="cuda", dtype=torch.int8)).permute(1, 0)
        out = ctx.empty((4096, 4096), np.int32)
        for _ in range(3):
            rt.MatMulInteger().run(ctx, a, b, out=out)
    ctx.sync()
print("done")
