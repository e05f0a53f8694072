# quick standalone probe of the CTA-pair path (run under `timeout`): small GEMM, forced plan, compare with numpy
import os, sys, numpy as np

os.environ.setdefault("RTEN_B200_F32_MODE", "tf32")  # these tools measure the single-pass TF32 kernels unless told otherwise
sys.path.insert(0, os.getcwd())
os.environ["RTEN_B200_FORCE_CTA2"] = "1"
os.environ["RTEN_B200_FORCE_BN"] = "128"
os.environ["RTEN_B200_FORCE_PAIR"] = "0"
os.environ["RTEN_B200_VERBOSE"] = "1"
import rten_b200 as rt
ctx = rt.Context(0)
rng = np.random.default_rng(0)
for (M, N, K) in [(256, 128, 64), (512, 256, 256), (1000, 384, 520)]:
    a = rng.standard_normal((M, K), dtype=np.float32)
    b = rng.standard_normal((K, N), dtype=np.float32)
    got = rt.MatMul().run(ctx, a, b).numpy()
    ref = a.astype(np.float64) @ b.astype(np.float64)
    print(M, N, K, "max err", float(np.abs(got - ref).max()), "ref scale", float(np.abs(ref).max()), flush=True)
