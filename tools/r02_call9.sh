#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA -k "attention or bert or tf32x3 or halo or gpt2_int8" > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "PASSED|FAILED|passed|failed" gpurun_out/c9_pytest.log | tail -12
grep -n "Error" gpurun_out/c9_pytest.log | head
timeout 600 python bench.py --model bert --steps 10 --warmup 3 --no-peaks --no-cpu-baseline --modes tf32 > gpurun_out/c9_bench_bert.json 2> gpurun_out/c9_bench_bert.err; echo "bench bert rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c9_bench_bert.json').read().strip().splitlines()[-1])
print('bert', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('top_kernels_us_per_step'))
PY
DECODE_MODES=fused timeout 300 python tools/decode_probe.py > gpurun_out/c9_decode.log 2>&1; tail -2 gpurun_out/c9_decode.log
