#!/bin/bash
# tools/build_variant.sh NAME [-DFLAG ...]: builds rmf_crowdsim_b200/_lib/variants/NAME.so with extra defines (A/B runs)
set -e
name=$1; shift
mkdir -p rmf_crowdsim_b200/_lib/variants
nvcc -std=c++17 -O3 --fmad=false -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math \
  -shared -cudart static "$@" -o rmf_crowdsim_b200/_lib/variants/$name.so rmf_crowdsim_b200/csrc/rcs.cu
