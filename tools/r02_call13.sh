#!/bin/bash
mkdir -p gpurun_out
PROBE=layers timeout 900 python tools/layer_probe.py > gpurun_out/c13_layer_probe.log 2>&1; echo "probe rc=$?"; tail -3 gpurun_out/c13_layer_probe.log
cp gpurun_out/layer_probe.txt gpurun_out/c13_layer_probe.txt
