#!/bin/bash
# ncu --set full captures of the shipped kernels on the BASELINE configs (one launch each: bounce 1 of a 4-sample
# batch at the config's full size; 32 samples on the Cornell box, whose traversal kernels are the one-ray-per-thread
# k_*_simple).  Usage (on the GPU box): tools/capture_traversal.sh <tag> c2 c3 c4 c5
# gpurun_out/ is capped at 64 MiB: only the raw and source CSV pages of each report are kept.
tag=$1; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  for k in k_extend k_shadow k_shade; do
    name=$k; iters=4
    if [ "$cfg" = "c2" ]; then iters=32; [ "$k" != "k_shade" ] && name=${k}_simple; fi
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^${name}\$" -s 0 -c 1 -f -o gpurun_out/${tag}_${cfg}_${k} \
      python tools/profile_step.py $cfg $iters > gpurun_out/${tag}_${cfg}_${k}.log 2>&1
    ncu -i gpurun_out/${tag}_${cfg}_${k}.ncu-rep --page raw --csv > gpurun_out/${tag}_${cfg}_${k}_raw.csv 2>/dev/null
    ncu -i gpurun_out/${tag}_${cfg}_${k}.ncu-rep --page source --csv > gpurun_out/${tag}_${cfg}_${k}_source.csv 2>/dev/null
    rm -f gpurun_out/${tag}_${cfg}_${k}.ncu-rep
  done
done
ls -la gpurun_out | grep ${tag}_ | head -60
