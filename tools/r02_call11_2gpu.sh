#!/bin/bash
# 2-GPU box: the whole GPU suite (the sharded test included), BERT with the merged QKV projection, weak-scaling lines
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c11_gpus.txt
timeout 1500 python -m pytest tests -m gpu -q -rA > gpurun_out/c11_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "FAILED|SKIPPED|passed|failed" gpurun_out/c11_pytest.log | tail -8
grep -n "Error" gpurun_out/c11_pytest.log | head
timeout 600 python bench.py --model bert --steps 10 --warmup 3 --no-peaks --no-cpu-baseline --modes tf32 > gpurun_out/c11_bench_bert.json 2> gpurun_out/c11_bench_bert.err; echo "bench bert rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c11_bench_bert.json').read().strip().splitlines()[-1])
print('bert', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('top_kernels_us_per_step'))
PY
for m in resnet50 bert resnet50_int8 gpt2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 10 --warmup 3 --model $m --no-peaks --no-extras --no-cpu-baseline > gpurun_out/c11_bench_${m}_n2.json 2> gpurun_out/c11_bench_${m}_n2.err; echo "$m n2 rc=$?"; tail -c 400 gpurun_out/c11_bench_${m}_n2.err | tail -3
  timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --model $m --no-peaks --no-extras --no-cpu-baseline > gpurun_out/c11_bench_${m}_n1.json 2> gpurun_out/c11_bench_${m}_n1.err; echo "$m n1 rc=$?"
done
python - <<'PY'
import json
for m in ["resnet50","bert","resnet50_int8","gpt2"]:
    for n in (1,2):
        try:
            d=json.loads(open(f'gpurun_out/c11_bench_{m}_n{n}.json').read().strip().splitlines()[-1])
            print(m, n, round(d['value'],1), d['unit'], 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'n_gpus', d['n_gpus'])
        except Exception as e:
            print(m, n, 'ERR', e)
PY
