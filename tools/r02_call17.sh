#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "halo_conv or conv_basic" > gpurun_out/c17_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "FAILED|passed|failed|Error|assert" gpurun_out/c17_pytest.log | tail -8
timeout 900 python tools/halo_sweep.py > gpurun_out/c17_halo_sweep.log 2>&1; echo "sweep rc=$?"; grep -E "==|model" gpurun_out/c17_halo_sweep.log
cp gpurun_out/halo_sweep.txt gpurun_out/c17_halo_sweep.txt
