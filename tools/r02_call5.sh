#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA -k "halo or conv or resnet50_model or mnist or tf32x3 or resnet50_b32" > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "PASSED|FAILED|passed|failed" gpurun_out/c5_pytest.log | tail -16
grep -n "AssertionError\|Error" gpurun_out/c5_pytest.log | head
timeout 900 python tools/halo_sweep.py > gpurun_out/c5_halo_sweep.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/halo_sweep.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-peaks --no-extras --no-cpu-baseline --modes tf32 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/c5_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c5_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('top_kernels_us_per_step'))
PY
