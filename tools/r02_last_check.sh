#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "FAILED|passed|failed" gpurun_out/l_pytest.log | tail -3
timeout 1200 python bench.py > gpurun_out/l_bench_default.json 2> gpurun_out/l_bench_default.err; echo "bench default rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/l_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/l_smoke.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/l_bench_default.json').read().strip().splitlines()[-1])
print('default:', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'layerwise', round(d['roofline']['layerwise']['frac'],3), 'launches', d['gpu_launches'])
print('x3:', round(d['modes']['tf32x3']['value'],1))
print({k:(round(v,1) if isinstance(v,float) else v) for k,v in d['also'].items() if k!='clocks' and k!='peaks'})
print('cpu', d['cpu_baseline'])
PY
