#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29617 tools/sharded_int8_check.py > gpurun_out/c28_sharded4.log 2>&1; echo "sharded4 rc=$?"; grep -E "world|exchange" gpurun_out/c28_sharded4.log
for m in resnet50_int8 resnet50; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 4 --steps 10 --warmup 3 --model $m --no-peaks --no-extras --no-cpu-baseline > gpurun_out/c28_bench_${m}_n4.json 2> gpurun_out/c28_bench_${m}_n4.err; echo "$m n4 rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c28_bench_${m}_n4.json').read().strip().splitlines()[-1])
print('$m gpus 4', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), d.get('comm'), d.get('cuda_graph'))
PY
done
