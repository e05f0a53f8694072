"""Summarise an .ncu-rep (ncu --set full) into a small CSV + JSON under profiles/: per launch duration, DRAM bytes,
DRAM / L2 / tensor-pipe utilisation, registers.  Usage: python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/NAME"""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = {
    "Kernel Name": "kernel", "Grid Size": "grid", "gpu__time_duration.sum": "duration_ns",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "launch__registers_per_thread": "regs", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
}
cols = [(k, v) for k, v in want.items() if k in idx]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_ns(val, unit):
    v = float(val.replace(",", ""))
    return v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)


recs = []
for d in data:
    r = {}
    for k, name in cols:
        val, unit = d[idx[k]], units[idx[k]]
        if name in ("dram_read", "dram_write"):
            r[name] = to_bytes(val, unit)
        elif name == "duration_ns":
            r[name] = to_ns(val, unit)
        elif name in ("kernel", "grid"):
            r[name] = val[:60]
        else:
            try:
                r[name] = float(val.replace(",", ""))
            except ValueError:
                r[name] = val
    recs.append(r)
with open(out + ".csv", "w", newline="") as f:
    w = csv.DictWriter(f, fieldnames=[n for _, n in cols])
    w.writeheader()
    w.writerows(recs)
umma = [r for r in recs if "umma_gemm" in r.get("kernel", "") or "umma_halo" in r.get("kernel", "")]
tot = sum(r["duration_ns"] for r in recs)


def agg(rs):
    t = sum(r["duration_ns"] for r in rs)
    o = {"launches": len(rs), "time_us": t / 1e3, "share_of_time": t / tot if tot else None,
         "avg_dram_bytes_per_launch": sum(r.get("dram_read", 0) + r.get("dram_write", 0) for r in rs) / len(rs),
         "achieved_dram_gbs": sum(r.get("dram_read", 0) + r.get("dram_write", 0) for r in rs) / t if t else None}
    for key in ("tensor_pipe_pct", "dram_pct", "l2_pct", "sm_pct"):
        if rs and key in rs[0] and isinstance(rs[0][key], float):
            o["time_weighted_" + key] = sum(r[key] * r["duration_ns"] for r in rs) / t
    return o


by = {}
for r in recs:
    name = r.get("kernel", "?").split("(")[0].replace("void ", "").strip()[:48]
    by.setdefault(name, []).append(r)
summ = {
    "launches": len(recs), "umma_launches": len(umma), "total_us": tot / 1e3,
    "umma_share_of_time": sum(r["duration_ns"] for r in umma) / tot if tot else None,
    "umma_avg_dram_bytes_per_launch": sum(r["dram_read"] + r["dram_write"] for r in umma) / len(umma) if umma else None,
    "umma_avg_duration_us": sum(r["duration_ns"] for r in umma) / len(umma) / 1e3 if umma else None,
    "umma_time_weighted_tensor_pipe_pct": sum(r["tensor_pipe_pct"] * r["duration_ns"] for r in umma) / sum(r["duration_ns"] for r in umma) if umma else None,
    "umma_time_weighted_dram_pct": sum(r["dram_pct"] * r["duration_ns"] for r in umma) / sum(r["duration_ns"] for r in umma) if umma and "dram_pct" in umma[0] else None,
    "total_dram_bytes": sum(r.get("dram_read", 0) + r.get("dram_write", 0) for r in recs),
    "by_kernel": {k: agg(v) for k, v in sorted(by.items(), key=lambda kv: -sum(r["duration_ns"] for r in kv[1]))},
}
json.dump(summ, open(out + ".json", "w"), indent=1)
print(json.dumps({k: v for k, v in summ.items() if k != "by_kernel"}, indent=1))
