#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rA > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "PASSED|FAILED|passed|failed" gpurun_out/c2_pytest.log | tail -45
timeout 300 python tools/decode_probe.py > gpurun_out/c2_decode.log 2>&1; echo "decode rc=$?"; cat gpurun_out/c2_decode.log | tail -5
timeout 600 ncu --set full --clock-control none -k regex:'softmax|layer_norm' -o gpurun_out/c2_rowops python tools/rowops_target.py > gpurun_out/c2_ncu_rowops.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/c2_rowops.ncu-rep gpurun_out/c2_rowops > /dev/null 2>&1; cat gpurun_out/c2_rowops.csv
DECODE_STEPS=3 DECODE_MODES=fused timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c2_decode_launches.csv python tools/decode_probe.py > gpurun_out/c2_decode_ncu.log 2>&1; echo "ncu decode rc=$?"
tail -80 gpurun_out/c2_decode_launches.csv | cut -c1-200
