"""Run each GPU parity check in its own subprocess with a timeout (a hung kernel must not take the whole
GPU lease with it) and write a one-line verdict per check to gpurun_out/probe.log."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_one(name):
    import gpu_checks
    import rten_b200 as rt
    from oracle import oracle
    fn = dict(gpu_checks.ALL_CHECKS)[name]
    print(fn(rt, oracle))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_one(sys.argv[2])
        sys.exit(0)
    import gpu_checks
    names = sys.argv[1:] or [n for n, _ in gpu_checks.ALL_CHECKS]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "probe.log"), "a")
    for n in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "--one", n], capture_output=True, text=True, timeout=240)
            tail = (p.stdout.strip().splitlines() or [""])[-1] if p.returncode == 0 else (p.stderr.strip().splitlines() or ["?"])[-1]
            verdict = "PASS" if p.returncode == 0 else "FAIL"
            if p.returncode != 0:
                sys.stderr.write(p.stderr[-3000:] + "\n")
        except subprocess.TimeoutExpired:
            verdict, tail = "TIMEOUT", ""
        line = f"{verdict:8s} {n:16s} {time.time() - t0:6.1f}s  {tail}"
        print(line, flush=True)
        log.write(line + "\n")
        log.flush()
