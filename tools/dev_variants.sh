#!/bin/bash
# Kernel experiments without touching the product library: builds build/libmsda_<tag>.so from
# csrc/fused.cu compiled with extra flags (bf16 / TPH=4 instances only, FUSED_DEV_ONLY) linked with
# the other objects of the last full build.  Use with APOLLO_B200_LIB=build/libmsda_<tag>.so.
#   tools/dev_variants.sh tag1 "-DFOO=1" tag2 "-DFOO=2" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p build
C=apollo-vision-net_b200/csrc
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude"
for f in abi msda_fwd msda_bwd point_sampling rowops bev_prep wgrad coarse_scatter; do       # cached objects of the other sources
  if [ ! -f build/$f.o ] || [ $C/$f.cu -nt build/$f.o ] || [ include/msda_b200.h -nt build/$f.o ]; then
    $NV -c $C/$f.cu -o build/$f.o &
  fi
done
wait
while [ $# -ge 2 ]; do
  tag=$1; flags=$2; shift 2
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DFUSED_DEV_ONLY $flags \
      -Iinclude -c $C/fused.cu -o build/fused_$tag.o &&
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/libmsda_$tag.so \
      build/abi.o build/msda_fwd.o build/msda_bwd.o build/point_sampling.o build/rowops.o build/bev_prep.o build/wgrad.o build/coarse_scatter.o build/fused_$tag.o &&
    echo "built build/libmsda_$tag.so" ) &
done
wait
