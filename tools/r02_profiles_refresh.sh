#!/bin/bash
# refresh the captures whose kernels changed after tools/r02_profiles.sh ran (integer plain epilogue, Gelu epilogue, attention)
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
for spec in "resnet50_int8 tf32" "bert tf32" "gpt2 tf32"; do
  set -- $spec; m=$1; mode=$2; tag=$m
  plans=gpurun_out/r02_plans_$tag.txt
  timeout 300 python tools/profile_target.py --model $m --mode $mode --plans $plans > /dev/null 2> /dev/null
  timeout 400 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_$tag.csv python tools/profile_target.py --model $m --mode $mode --plans $plans > /dev/null 2> gpurun_out/r02_verbose_$tag.log; echo "$tag launches rc=$?"
  python tools/layer_table.py gpurun_out/r02_launches_$tag.csv gpurun_out/r02_verbose_$tag.log > gpurun_out/r02_layers_$tag.txt; tail -7 gpurun_out/r02_layers_$tag.txt
done
timeout 420 $NCU --set full -f -o /tmp/r02_resnet50_int8 python tools/profile_target.py --model resnet50_int8 --plans gpurun_out/r02_plans_resnet50_int8.txt > /dev/null 2> /dev/null; echo "int8 full rc=$?"
python tools/ncu_summary.py /tmp/r02_resnet50_int8.ncu-rep gpurun_out/r02_ncu_resnet50_int8 > gpurun_out/r02_ncu_resnet50_int8_summary.txt 2>&1; head -12 gpurun_out/r02_ncu_resnet50_int8_summary.txt
