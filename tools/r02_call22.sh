#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "plans or matmul or conv or bert or resnet50_b32 or resnet50_model or sequence or gemm_op or tf32x3" > gpurun_out/c22_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "FAILED|passed|failed|Error" gpurun_out/c22_pytest.log | tail -5
for m in resnet50 bert; do
timeout 600 python bench.py --model $m --steps 20 --warmup 5 --no-peaks --no-extras --no-cpu-baseline > gpurun_out/c22_bench_$m.json 2> gpurun_out/c22_bench_$m.err; echo "bench $m rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c22_bench_$m.json').read().strip().splitlines()[-1])
print('$m', round(d['value'],1), round(d['ms_per_step'],4), 'x3', round(d['modes']['tf32x3']['value'],1), d.get('top_kernels_us_per_step'))
PY
done
