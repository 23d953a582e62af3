#!/bin/bash
# tools/ab.sh SPP V1 V2 ...  -> K1 timing of build/variants/libuecrt_V.so on the final scene, same box, same call
spp=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v" | tee -a gpurun_out/ab.log
  UECRT_LIB=build/variants/libuecrt_$v.so timeout 300 python tools/tc_perf.py $spp tensor 3 2>&1 | tee -a gpurun_out/ab.log
done
