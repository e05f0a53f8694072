#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA -k "attention or bert or tf32x3" > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "PASSED|FAILED|passed|failed" gpurun_out/c7_pytest.log | tail -12
grep -n "Error" gpurun_out/c7_pytest.log | head
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'umma_' -o gpurun_out/c7_halo python tools/halo_ncu_target.py > gpurun_out/c7_halo_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/c7_halo.ncu-rep gpurun_out/c7_halo > /dev/null 2>&1; cat gpurun_out/c7_halo.csv | cut -c1-220
timeout 600 python bench.py --model bert --steps 10 --warmup 3 --no-peaks --no-cpu-baseline > gpurun_out/c7_bench_bert.json 2> gpurun_out/c7_bench_bert.err; echo "bench bert rc=$?"; tail -c 300 gpurun_out/c7_bench_bert.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c7_bench_bert.json').read().strip().splitlines()[-1])
print('bert', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('top_kernels_us_per_step'))
print('bert x3', d['modes']['tf32x3']['value'], d['modes']['tf32x3']['ms_per_step'], d['modes']['tf32x3'].get('top_kernels_us_per_step'))
PY
