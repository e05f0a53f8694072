#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tf32x3 or reference_rule or mnist or resnet50_b32 or bert_b16" > gpurun_out/c29_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "FAILED|passed|failed" gpurun_out/c29_pytest.log | tail -3
for m in resnet50 bert; do
timeout 600 python bench.py --model $m --steps 10 --warmup 3 --no-peaks --no-extras --no-cpu-baseline --modes tf32x3 > gpurun_out/c29_bench_${m}_x3.json 2> gpurun_out/c29_bench_${m}_x3.err; echo "x3 $m rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c29_bench_${m}_x3.json').read().strip().splitlines()[-1])
print('$m x3', round(d['value'],1), round(d['ms_per_step'],4), d.get('top_kernels_us_per_step'))
PY
done
