"""Merge an ncu launch list (gpu__time_duration per launch) of ONE model pass with the host's per-launch tile
configuration (RTEN_B200_VERBOSE lines, same order).  Usage: merge_launches.py launches.csv verbose.log [last_n]"""
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if "umma_gemm" in r["Kernel Name"]]
cfg = [l.strip()[12:] for l in open(sys.argv[2]) if l.startswith("[umma_gemm]")]
n = min(len(rows), len(cfg))
if len(sys.argv) > 3:  # only the last N launches (one model pass)
    n = min(n, int(sys.argv[3]))
rows, cfg = rows[-n:], cfg[-n:]
tot = 0.0
out = []
for r, c in zip(rows, cfg):
    us = float(r["Metric Value"].replace(",", "")) / (1e3 if r["Metric Unit"] == "ns" else 1.0)
    tot += us
    kv = dict(x.split("=") for x in c.split() if "=" in x)
    M, N, K = int(kv["M"]), int(kv["N"]), int(kv["K"])
    fl = 2.0 * M * N * K
    if kv.get("conv") == "1" and int(kv["kb"]) * 32 != K and K % 32 == 0:
        pass
    out.append((us, fl / us / 1e6, c))
for us, tf, c in out:
    print(f"{us:7.1f} us {tf:6.1f} TF/s  {c}")
print(f"total umma {tot:.1f} us over {n} launches")
