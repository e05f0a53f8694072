"""What the GPU box's host CPU really offers: affinity, cgroup quota, and oracle throughput vs thread count."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
print("affinity", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count())
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try:
        print(p, open(p).read().strip())
    except Exception as e:
        print(p, "-", type(e).__name__)
os.system("grep -m1 'model name' /proc/cpuinfo; grep -c processor /proc/cpuinfo; cat /proc/loadavg")
from oracle import oracle
import numpy as np, model_ref, bench
spec = bench.make_spec(oracle, "resnet50")
x = bench.make_inputs(oracle, "resnet50", 32)["x"]
for nt in (8, 16, 32, 64, 128):
    oracle.lib().rto_set_num_threads(nt)
    ar = oracle.Arena()
    model_ref.resnet50_oracle(oracle, spec, x, ar)
    t = time.perf_counter(); model_ref.resnet50_oracle(oracle, spec, x, ar); dt = time.perf_counter() - t
    print(f"threads {nt:4d}: {32 / dt:7.1f} img/s", flush=True)
# single GEMM scaling
a = np.random.rand(2048, 2048).astype(np.float32); b = np.random.rand(2048, 2048).astype(np.float32)
for nt in (1, 8, 32, 128):
    oracle.lib().rto_set_num_threads(nt)
    oracle.gemm_f32(a, b)
    t = time.perf_counter(); oracle.gemm_f32(a, b); dt = time.perf_counter() - t
    print(f"gemm 2048^3 threads {nt:4d}: {2 * 2048**3 / dt / 1e9:8.1f} GFLOP/s", flush=True)
