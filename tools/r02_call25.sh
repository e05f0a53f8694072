#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dql or int8 or integer or gpt2" > gpurun_out/c25_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "FAILED|passed|failed|Error|assert" gpurun_out/c25_pytest.log | tail -8
timeout 600 python bench.py --model resnet50_int8 --steps 10 --warmup 3 --no-peaks --no-extras --no-cpu-baseline > gpurun_out/c25_bench_int8.json 2> gpurun_out/c25_bench_int8.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c25_bench_int8.json').read().strip().splitlines()[-1])
print('int8', round(d['value'],1), round(d['ms_per_step'],4), d.get('top_kernels_us_per_step'))
PY
