#!/bin/bash
# Row-kernel experiments without touching the product library: builds build/libmsda_<tag>.so from
# csrc/rowops.cu compiled with extra flags, linked with cached objects of the other sources.
# Use with APOLLO_B200_LIB=build/libmsda_<tag>.so (tools/rowops_bench.py, bench.py).
#   tools/rowops_variants.sh tag1 "-DLN_FWD_DEPTH=1" tag2 "-DLN_BWD_MINBLOCKS=3" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p build gpurun_out/objcache
C=apollo-vision-net_b200/csrc
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude"
OTHERS="abi msda_fwd msda_bwd point_sampling bev_prep wgrad coarse_scatter fused mha dcnv3"
for f in $OTHERS; do
  if [ ! -f gpurun_out/objcache/$f.o ] || [ $C/$f.cu -nt gpurun_out/objcache/$f.o ] || [ include/msda_b200.h -nt gpurun_out/objcache/$f.o ]; then
    $NV -c $C/$f.cu -o gpurun_out/objcache/$f.o &
  fi
done
wait
objs=""
for f in $OTHERS; do objs="$objs gpurun_out/objcache/$f.o"; done
while [ $# -ge 2 ]; do
  tag=$1; flags=$2; shift 2
  ( $NV $flags -c $C/rowops.cu -o gpurun_out/objcache/rowops_$tag.o &&
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/libmsda_$tag.so $objs gpurun_out/objcache/rowops_$tag.o &&
    echo "built build/libmsda_$tag.so" ) &
done
wait
