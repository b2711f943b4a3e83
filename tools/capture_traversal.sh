#!/bin/bash
# ncu --set full captures of the shipped traversal kernels on the BASELINE configs (one launch each: the depth-1
# bounce of a 4-sample batch).  Usage (on the GPU box): tools/capture_traversal.sh <tag> c3 c4 c5
tag=$1; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  for k in k_extend k_shadow k_shade; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^${k}\$|${k}<" -s 0 -c 1 -f -o gpurun_out/${tag}_${cfg}_${k} \
      python tools/profile_step.py $cfg 4 > gpurun_out/${tag}_${cfg}_${k}.log 2>&1
    ncu -i gpurun_out/${tag}_${cfg}_${k}.ncu-rep --page raw --csv > gpurun_out/${tag}_${cfg}_${k}_raw.csv 2>/dev/null
    ncu -i gpurun_out/${tag}_${cfg}_${k}.ncu-rep --page source --csv > gpurun_out/${tag}_${cfg}_${k}_source.csv 2>/dev/null
    rm -f gpurun_out/${tag}_${cfg}_${k}.ncu-rep  # gpurun_out/ is capped at 64 MiB: the two CSV pages are what gets read
  done
done
ls -la gpurun_out | grep ${tag}_ | head -40
