#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -rA -s > gpurun_out/c27_pytest_sharded.log 2>&1; echo "sharded rc=$?"; grep -E "exchange through|world 2|passed|failed" gpurun_out/c27_pytest_sharded.log | tail -5
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dql or int8" > gpurun_out/c27_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "FAILED|passed|failed" gpurun_out/c27_pytest.log | tail -3
for n in 2 1; do
  if [ $n = 2 ]; then L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"; else L="python"; fi
  timeout 900 $L bench.py --gpus $n --steps 10 --warmup 3 --model resnet50_int8 --no-peaks --no-extras --no-cpu-baseline > gpurun_out/c27_bench_int8_n$n.json 2> gpurun_out/c27_bench_int8_n$n.err; echo "int8 n$n rc=$?"
done
python - <<'PY'
import json
v={}
for n in (1,2):
    d=json.loads(open(f'gpurun_out/c27_bench_int8_n{n}.json').read().strip().splitlines()[-1]); v[n]=d
    print('int8 gpus', n, round(d['value'],1), 'ms/step', round(d['ms_per_step'],4))
print('efficiency', v[2]['value']/(2*v[1]['value']))
PY
