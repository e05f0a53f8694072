#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -rA -s > gpurun_out/c16_pytest_sharded.log 2>&1; echo "sharded rc=$?"; grep -E "exchange through|world 2|passed|failed" gpurun_out/c16_pytest_sharded.log | tail -5
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tf32x3 or mnist or reference_rule or conv_basic or resnet50_b32" > gpurun_out/c16_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "FAILED|passed|failed|Error" gpurun_out/c16_pytest.log | tail -5
for n in 2 1; do
  if [ $n = 2 ]; then L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"; else L="python"; fi
  timeout 900 $L bench.py --gpus $n --steps 10 --warmup 3 --model resnet50_int8 --no-peaks --no-extras --no-cpu-baseline > gpurun_out/c16_bench_int8_n$n.json 2> gpurun_out/c16_bench_int8_n$n.err; echo "int8 n$n rc=$?"
done
timeout 600 python bench.py --model resnet50 --steps 10 --warmup 3 --no-peaks --no-extras --no-cpu-baseline --modes tf32x3 > gpurun_out/c16_bench_resnet50_x3.json 2> gpurun_out/c16_bench_resnet50_x3.err; echo "x3 rc=$?"
python - <<'PY'
import json
for f in ("int8_n1","int8_n2","resnet50_x3"):
    try:
        d=json.loads(open(f'gpurun_out/c16_bench_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value'],1), d['unit'], 'ms/step', round(d['ms_per_step'],4), d.get('top_kernels_us_per_step'))
    except Exception as e:
        print(f, 'ERR', e)
PY
