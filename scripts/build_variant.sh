#!/bin/bash
# Builds deodr_b200/libdeodr_b200_<name>.so with extra nvcc flags (A/B experiments: DEODR_B200_LIB selects it at run time).
#   scripts/build_variant.sh tilez4 -DDEODR_TILEZ_MIN_CTAS=4
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build_$name
pids=()
for u in kernels.cu kernels_bwd.cu scene_ops.cu host_api.cu host_simd.cpp; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-pthread "$@" \
     -c -o build_$name/${u%.*}.o deodr_b200/csrc/$u &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p || exit 1; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC,-pthread -o deodr_b200/libdeodr_b200_$name.so build_$name/*.o
ls -la deodr_b200/libdeodr_b200_$name.so
