#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "glue or bert or gpt2 or generator or model_executor or mnist" > gpurun_out/c31_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "FAILED|passed|failed" gpurun_out/c31_pytest.log | tail -3
timeout 600 python bench.py --model bert --steps 20 --warmup 5 --no-peaks --no-extras --no-cpu-baseline --modes tf32 > gpurun_out/c31_bench_bert.json 2> gpurun_out/c31_bench_bert.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c31_bench_bert.json').read().strip().splitlines()[-1])
print('bert', round(d['value'],1), round(d['ms_per_step'],4), d.get('top_kernels_us_per_step'))
PY
