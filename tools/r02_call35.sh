#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gpt2 or generator" > gpurun_out/c35_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "FAILED|passed|failed|assert|Error" gpurun_out/c35_pytest.log | tail -4
timeout 600 python bench.py --model gpt2 --steps 20 --warmup 5 --no-peaks --no-extras --no-cpu-baseline > gpurun_out/c35_bench_gpt2.json 2> gpurun_out/c35_bench_gpt2.err; echo "bench rc=$?"; tail -3 gpurun_out/c35_bench_gpt2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/c35_bench_gpt2.json').read().strip().splitlines()[-1])
print('gpt2', round(d['value'],1), round(d['ms_per_step'],4), 'prefill tok/s', round(d.get('prefill_tokens_per_sec',0)))
PY
