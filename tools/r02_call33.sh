#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "attention_decode or gpt2 or generator or quantized" > gpurun_out/c33_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "FAILED|passed|failed|assert" gpurun_out/c33_pytest.log | tail -3
timeout 600 python bench.py --model gpt2 --steps 20 --warmup 5 --no-peaks --no-extras --no-cpu-baseline > gpurun_out/c33_bench_gpt2.json 2> gpurun_out/c33_bench_gpt2.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c33_bench_gpt2.json').read().strip().splitlines()[-1])
print('gpt2', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d.get('top_kernels_us_per_step'))
PY
