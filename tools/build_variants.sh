#!/bin/bash
# Tuning aid: build libattpc_b200.so variants with other finalize tilings into build/variants/ (git-ignored; travels
# with gpurun).  usage: tools/build_variants.sh "512 5120" "256 3584" ...
set -eu
mkdir -p build/variants
for v in "$@"; do
  set -- $v
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
    -DATTPC_FIN_THREADS=$1 -DATTPC_FIN_ITEMS=$2 ${3:+$3} \
    -o build/variants/lib_t$1_i$2${3:+_c${3##*=}}.so attpc_engine_b200/csrc/attpc_b200.cu &
done
wait
ls -l build/variants
