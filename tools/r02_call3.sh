#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rA -k "quantized_linear or attention_decode or skinny_f32 or gpt2 or plans or matmul or gemm or softmax or layer_norm" > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "PASSED|FAILED|passed|failed" gpurun_out/c3_pytest.log | tail -25
grep -n "AssertionError" gpurun_out/c3_pytest.log | head
timeout 300 python tools/decode_probe.py > gpurun_out/c3_decode.log 2>&1; echo "decode rc=$?"; tail -5 gpurun_out/c3_decode.log
DECODE_STEPS=3 DECODE_MODES=fused timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c3_decode_launches.csv python tools/decode_probe.py > gpurun_out/c3_decode_ncu.log 2>&1; echo "ncu decode rc=$?"
DECODE_STEPS=2 DECODE_MODES=fused timeout 600 ncu --set full --import-source on --clock-control none -k regex:'qlinear|attn_decode' -s 130 -c 8 -o gpurun_out/c3_decode_full python tools/decode_probe.py > gpurun_out/c3_decode_full.log 2>&1; echo "ncu full rc=$?"
