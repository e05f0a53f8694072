#!/bin/bash
# GPU call 1 of round 2: descriptor probe, full GPU test suite, peaks + profiler check, plan sweep, row-kernel ncu.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 120 tools/bin/desc_probe > gpurun_out/c1_desc_probe.txt 2>&1; echo "desc_probe rc=$?"
timeout 900 python -m pytest tests -m gpu -q -rA > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/c1_pytest.log
timeout 300 python tools/peaks_probe.py > gpurun_out/c1_peaks.json 2> gpurun_out/c1_peaks.err; echo "peaks rc=$?"
timeout 900 python tools/plan_sweep.py > gpurun_out/c1_plan_sweep.log 2>&1; echo "sweep rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'softmax|layer_norm|unary|dql|minmax' -o gpurun_out/c1_rowops python tools/rowops_target.py > gpurun_out/c1_ncu_rowops.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/c1_rowops.ncu-rep gpurun_out/c1_rowops > gpurun_out/c1_rowops_summary.txt 2>&1
ls -la gpurun_out | tail -20
