#!/bin/bash
# Build an experimental variant of the library: tools/build_variant.sh <name> <extra nvcc -D flags...>
# -> gpurun_variants/lib_<name>.so ; run with STRELKA_B200_LIB=gpurun_variants/lib_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p gpurun_variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false --extended-lambda \
  --expt-relaxed-constexpr -Xptxas -v -Xcompiler -fPIC,-Wno-unknown-pragmas "$@" -shared -o gpurun_variants/lib_$name.so strelka_b200/csrc/unity.cu -lcudart 2>&1 > gpurun_variants/_$name.log 2>&1 || (grep error gpurun_variants/_$name.log; false)
ls -la gpurun_variants/lib_$name.so
