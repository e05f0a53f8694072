#!/bin/bash
# Round-end validation on a 2-GPU box: whole GPU suite, default bench line, reference arm, weak-scaling lines, smoke
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/f_gpus.txt
timeout 1800 python -m pytest tests -m gpu -q -rA > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "FAILED|SKIPPED|passed|failed" gpurun_out/f_pytest.log | tail -6
timeout 1200 python bench.py > gpurun_out/f_bench_default.json 2> gpurun_out/f_bench_default.err; echo "bench default rc=$?"; tail -c 300 gpurun_out/f_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "bench ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/f_smoke.log
for m in resnet50 bert resnet50_int8 gpt2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 10 --warmup 3 --model $m --no-peaks --no-extras --no-cpu-baseline > gpurun_out/f_bench_${m}_n2.json 2> gpurun_out/f_bench_${m}_n2.err; echo "$m n2 rc=$?"
  timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --model $m --no-peaks --no-extras --no-cpu-baseline > gpurun_out/f_bench_${m}_n1.json 2> gpurun_out/f_bench_${m}_n1.err; echo "$m n1 rc=$?"
done
python - <<'PY' | tee gpurun_out/f_multi_gpu.txt
import json
for m in ["resnet50","bert","resnet50_int8","gpt2"]:
    v={}
    for n in (1,2):
        try:
            d=json.loads(open(f'gpurun_out/f_bench_{m}_n{n}.json').read().strip().splitlines()[-1])
            v[n]=d
            print(m, 'gpus', n, round(d['value'],1), d['unit'], 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))
        except Exception as e:
            print(m, n, 'ERR', e)
    if 1 in v and 2 in v:
        print(f"   weak-scaling efficiency at 2 GPUs: {v[2]['value'] / (2 * v[1]['value']) * 100:.1f} %")
PY
python - <<'PY'
import json
d=json.loads(open('gpurun_out/f_bench_default.json').read().strip().splitlines()[-1])
print('default:', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'layerwise', d['roofline']['layerwise']['frac'])
print('x3:', d['modes']['tf32x3']['value'])
print('also:', {k:v for k,v in d['also'].items() if k!='clocks'})
PY
