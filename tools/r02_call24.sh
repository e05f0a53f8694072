#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gelu or unary or bert or matmul or gpt2" > gpurun_out/c24_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "FAILED|passed|failed|Error|assert" gpurun_out/c24_pytest.log | tail -8
timeout 600 python bench.py --model bert --steps 20 --warmup 5 --no-peaks --no-extras --no-cpu-baseline > gpurun_out/c24_bench_bert.json 2> gpurun_out/c24_bench_bert.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c24_bench_bert.json').read().strip().splitlines()[-1])
print('bert', round(d['value'],1), round(d['ms_per_step'],4), 'x3', round(d['modes']['tf32x3']['value'],1), d.get('top_kernels_us_per_step'))
PY
