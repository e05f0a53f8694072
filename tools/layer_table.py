"""Join an ncu launch list (gpu__time_duration.sum per launch, `--csv --log-file`) of ONE profiled pass with the launch
plans the library printed for the tensor-core launches of that pass (RTEN_B200_VERBOSE lines after the
"== profiled pass ==" marker of tools/profile_target.py, same order).  Prints every launch, per-kernel totals and, for
the tensor-core launches, TFLOP/s and CTAs at work.
Usage: layer_table.py launches.csv verbose.log > profiles/r02_layers_<model>.txt"""
import csv
import re
import sys
from collections import OrderedDict

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
log = open(sys.argv[2]).read()
log = log.split("== profiled pass ==")[-1]
plans = [l.strip() for l in log.splitlines() if l.startswith("[umma_gemm]") or l.startswith("[umma_halo]")]
pi = 0
tot = 0.0
by_kernel = OrderedDict()
print("#     us   TF/s  CTAs  kernel / plan")
for r in rows:
    name = r["Kernel Name"]
    us = float(r["Metric Value"].replace(",", "")) / (1e3 if r["Metric Unit"] in ("ns", "nsecond") else 1.0)
    tot += us
    short = re.sub(r"\(.*", "", name)[:70]
    k = by_kernel.setdefault(short, [0, 0.0])
    k[0] += 1
    k[1] += us
    extra = ""
    if ("umma_gemm_kernel" in name or "umma_halo_kernel" in name) and pi < len(plans):
        c = plans[pi]
        pi += 1
        kv = dict(x.split("=") for x in c.replace(":", " ").split() if "=" in x and x.count("=") == 1)
        try:
            if c.startswith("[umma_gemm]"):
                fl = 2.0 * int(kv["M"]) * int(kv["N"]) * int(kv["K"])
                ctas = min(148, int(kv["units"]) * (2 if kv.get("cta2") == "1" else 1))
            else:
                m = re.search(r"B=(\d+) (\d+)x(\d+) C=(\d+) N=(\d+) k=(\d+)x(\d+)", c)
                b, oh, ow, ci, n, kh, kw = (int(v) for v in m.groups())
                fl = 2.0 * b * oh * ow * n * ci * kh * kw
                ctas = min(148, int(kv["units"]))
            extra = f"{fl / us / 1e6:6.1f} {ctas:5d}  {c}"
        except Exception:
            extra = f"     ?     ?  {c}"
    else:
        extra = f"     -     -  {short}"
    print(f"{us:8.1f} {extra}")
print(f"# total {tot:.1f} us over {len(rows)} launches (cold-cache, serialised by the profiler); {pi} tensor-core launches joined with {len(plans)} plan lines")
for k, (n, us) in sorted(by_kernel.items(), key=lambda kv: -kv[1][1]):
    print(f"# {us:9.1f} us {100 * us / tot:5.1f} %  {n:4d}x  {k}")
