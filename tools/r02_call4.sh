#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA -k "quantized_linear or attention_decode or gpt2 or dql" > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "PASSED|FAILED|passed|failed" gpurun_out/c4_pytest.log | tail -12
grep -n "AssertionError\|Error" gpurun_out/c4_pytest.log | head
DECODE_MODES=fused timeout 300 python tools/decode_probe.py > gpurun_out/c4_decode.log 2>&1; echo "decode rc=$?"; tail -3 gpurun_out/c4_decode.log
DECODE_STEPS=3 DECODE_MODES=fused timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c4_decode_launches.csv python tools/decode_probe.py > gpurun_out/c4_decode_ncu.log 2>&1; echo "ncu decode rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/c4_bench.err
timeout 600 python bench.py --model gpt2 --steps 10 --warmup 3 --no-peaks > gpurun_out/c4_bench_gpt2.json 2> gpurun_out/c4_bench_gpt2.err; echo "bench gpt2 rc=$?"; tail -c 800 gpurun_out/c4_bench_gpt2.err
timeout 600 python bench.py --model resnet50_int8 --steps 10 --warmup 3 --no-peaks --no-cpu-baseline > gpurun_out/c4_bench_i8.json 2> gpurun_out/c4_bench_i8.err; echo "bench i8 rc=$?"; tail -c 800 gpurun_out/c4_bench_i8.err
timeout 600 python bench.py --model bert --steps 10 --warmup 3 --no-peaks --no-cpu-baseline > gpurun_out/c4_bench_bert.json 2> gpurun_out/c4_bench_bert.err; echo "bench bert rc=$?"; tail -c 800 gpurun_out/c4_bench_bert.err
