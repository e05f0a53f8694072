#!/bin/bash
# Round-2 ncu captures: per model a launch list (+ plan table) and an `ncu --set full` pass reduced on the box by
# tools/ncu_summary.py (the .ncu-rep files are too large to bring back; a few sample kernels are kept with source).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
NCU="ncu --clock-control none --profile-from-start off"
for spec in "resnet50 tf32" "resnet50 tf32x3" "bert tf32" "resnet50_int8 tf32" "gpt2 tf32"; do
  set -- $spec; m=$1; mode=$2; tag=$m; [ "$mode" = "tf32x3" ] && tag=${m}_x3
  plans=gpurun_out/r02_plans_$tag.txt
  timeout 600 python tools/profile_target.py --model $m --mode $mode --plans $plans > /dev/null 2> /dev/null   # measure plans once
  timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_$tag.csv python tools/profile_target.py --model $m --mode $mode --plans $plans > /dev/null 2> gpurun_out/r02_verbose_$tag.log; echo "$tag launches rc=$?"
  python tools/layer_table.py gpurun_out/r02_launches_$tag.csv gpurun_out/r02_verbose_$tag.log > gpurun_out/r02_layers_$tag.txt; tail -8 gpurun_out/r02_layers_$tag.txt
  timeout 1500 $NCU --set full -f -o /tmp/r02_$tag python tools/profile_target.py --model $m --mode $mode --plans $plans > /dev/null 2> /dev/null; echo "$tag full rc=$?"
  python tools/ncu_summary.py /tmp/r02_$tag.ncu-rep gpurun_out/r02_ncu_$tag > gpurun_out/r02_ncu_${tag}_summary.txt 2>&1; head -12 gpurun_out/r02_ncu_${tag}_summary.txt
done
# sample kernels with source: plain-epilogue conv, halo conv, fused attention, quantised linear, decode attention
timeout 900 $NCU --set full --import-source on -f -k regex:"umma_gemm_kernel|umma_halo" -c 6 -o gpurun_out/r02_ncu_samples_conv python tools/profile_target.py --model resnet50 --plans gpurun_out/r02_plans_resnet50.txt > /dev/null 2>&1; echo "samples conv rc=$?"
timeout 900 $NCU --set full --import-source on -f -k regex:"attn_fused|layer_norm_vec" -c 2 -o gpurun_out/r02_ncu_samples_bert python tools/profile_target.py --model bert --plans gpurun_out/r02_plans_bert.txt > /dev/null 2>&1; echo "samples bert rc=$?"
timeout 900 $NCU --set full --import-source on -f -k regex:"qlinear|attn_decode" -c 5 -o gpurun_out/r02_ncu_samples_decode python tools/profile_target.py --model gpt2 > /dev/null 2>&1; echo "samples decode rc=$?"
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
