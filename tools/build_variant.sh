#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG=.. ..."  -> build/variants/libuecrt_NAME.so  (tuning experiments)
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2
mkdir -p build/variants
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --extended-lambda -Xcompiler -fPIC -ccbin /usr/bin/g++"
$NV $flags -Xptxas -v -c uecraytracing_b200/csrc/kernels.cu -o build/variants/kernels_$name.o 2> build/variants/ptxas_$name.log
$NV $flags -c uecraytracing_b200/csrc/uecrt_api.cu -o build/variants/api_$name.o
[ -f build/host_api.o ] || make build/host_api.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o build/variants/libuecrt_$name.so build/variants/kernels_$name.o build/variants/api_$name.o build/host_api.o -lz -ldl
echo "$name: $(grep -A1 'k_render_pathsILb0' build/variants/ptxas_$name.log | grep -o 'Used [0-9]* registers') $(grep -A1 'k_render_pathsILb0' build/variants/ptxas_$name.log | grep -o '[0-9]* bytes spill stores')"
