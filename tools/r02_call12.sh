#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/layer_probe.py > gpurun_out/c12_layer_probe.log 2>&1; echo "probe rc=$?"; tail -5 gpurun_out/c12_layer_probe.log
for b in 8 16 64; do
  RTEN_BENCH_BATCH=$b timeout 600 python bench.py --model resnet50 --steps 20 --warmup 5 --no-peaks --no-extras --no-cpu-baseline --modes tf32 > gpurun_out/c12_bench_b$b.json 2> gpurun_out/c12_bench_b$b.err; echo "b$b rc=$?"
done
python - <<'PY'
import json
for b in (8,16,64):
    try:
        d=json.loads(open(f'gpurun_out/c12_bench_b{b}.json').read().strip().splitlines()[-1])
        print('batch', b, round(d['value'],1), 'img/s', round(d['ms_per_step'],4), 'ms')
    except Exception as e: print(b,'ERR',e)
PY
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -k "tf32x3 or reference_rule or resnet50_b32 or bert_b16 or mnist or matmul_shapes or conv_more or bert_model" > gpurun_out/c12_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "FAILED|passed|failed|Error" gpurun_out/c12_pytest.log | tail -8
for m in resnet50 bert; do
timeout 600 python bench.py --model $m --steps 10 --warmup 3 --no-peaks --no-extras --no-cpu-baseline --modes tf32x3 > gpurun_out/c12_bench_${m}_x3.json 2> gpurun_out/c12_bench_${m}_x3.err; echo "x3 $m rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c12_bench_${m}_x3.json').read().strip().splitlines()[-1])
print('$m x3', d['value'], d['ms_per_step'], d.get('top_kernels_us_per_step'))
PY
done
