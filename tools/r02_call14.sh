#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conv or matmul or gemm_op or plans or sequence or resnet50_model or bert_model or resnet50_b32 or bert_b16 or tf32x3" > gpurun_out/c14_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "FAILED|passed|failed|Error" gpurun_out/c14_pytest.log | tail -8
PROBE=layers timeout 900 python tools/layer_probe.py > gpurun_out/c14_layer_probe.log 2>&1; echo "probe rc=$?"; tail -1 gpurun_out/c14_layer_probe.log
cp gpurun_out/layer_probe.txt gpurun_out/c14_layer_probe.txt
for m in resnet50 bert; do
timeout 600 python bench.py --model $m --steps 20 --warmup 5 --no-peaks --no-extras --no-cpu-baseline --modes tf32 > gpurun_out/c14_bench_$m.json 2> gpurun_out/c14_bench_$m.err; echo "bench $m rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c14_bench_$m.json').read().strip().splitlines()[-1])
print('$m', d['value'], d['ms_per_step'], d['roofline'].get('layerwise'), d.get('top_kernels_us_per_step'))
PY
done
