#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "FAILED|passed|failed" gpurun_out/c10_pytest.log | tail -8
grep -n "Error" gpurun_out/c10_pytest.log | head
timeout 600 python bench.py --model bert --steps 10 --warmup 3 --no-peaks --no-cpu-baseline --modes tf32 > gpurun_out/c10_bench_bert.json 2> gpurun_out/c10_bench_bert.err; echo "bench bert rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c10_bench_bert.json').read().strip().splitlines()[-1])
print('bert', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('top_kernels_us_per_step'))
PY
timeout 900 python bench.py > gpurun_out/c10_bench_default.json 2> gpurun_out/c10_bench_default.err; echo "bench default rc=$?"; tail -c 300 gpurun_out/c10_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c10_bench_ref.json 2> gpurun_out/c10_bench_ref.err; echo "bench ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c10_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c10_smoke.log
