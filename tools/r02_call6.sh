#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA -k "halo or model_executor or generator or conv_basic or resnet50_model or resnet50_b32 or mnist" > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "PASSED|FAILED|passed|failed" gpurun_out/c6_pytest.log | tail -12
grep -n "Error" gpurun_out/c6_pytest.log | head
timeout 900 python tools/halo_sweep.py > gpurun_out/c6_halo_sweep.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/halo_sweep.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-peaks --no-extras --no-cpu-baseline --modes tf32 > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/c6_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c6_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('top_kernels_us_per_step'))
PY
