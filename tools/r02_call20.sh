#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/dual_stream_probe.py > gpurun_out/c20_dual.log 2>&1; echo "dual rc=$?"; tail -5 gpurun_out/c20_dual.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "halo_conv or conv_basic or resnet50_model or resnet50_b32 or plans" > gpurun_out/c20_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "FAILED|passed|failed|Error" gpurun_out/c20_pytest.log | tail -5
RTEN_B200_VERBOSE=1 timeout 600 python bench.py --model resnet50 --steps 20 --warmup 5 --no-peaks --no-extras --no-cpu-baseline --modes tf32 > gpurun_out/c20_bench_resnet50.json 2> gpurun_out/c20_bench_resnet50.err; echo "bench rc=$?"
grep -c "autotune\] halo" gpurun_out/c20_bench_resnet50.err; grep "umma_halo\]" gpurun_out/c20_bench_resnet50.err | sort | uniq -c | head
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c20_bench_resnet50.json').read().strip().splitlines()[-1])
print('resnet50', d['value'], d['ms_per_step'], d.get('top_kernels_us_per_step'))
PY
