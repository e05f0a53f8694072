#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/c30_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "FAILED|passed|failed" gpurun_out/c30_pytest.log | tail -3
for m in resnet50 bert resnet50_int8; do
timeout 600 python bench.py --model $m --steps 20 --warmup 5 --no-peaks --no-extras --no-cpu-baseline --modes $( [ $m = resnet50_int8 ] && echo int8 || echo tf32 ) > gpurun_out/c30_bench_$m.json 2> gpurun_out/c30_bench_$m.err; echo "bench $m rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c30_bench_$m.json').read().strip().splitlines()[-1])
print('$m', round(d['value'],1), round(d['ms_per_step'],4))
PY
done
