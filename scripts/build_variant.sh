#!/bin/bash
# scripts/build_variant.sh name [-D...]: build variants/name.so and print the registers of the fused kernels
name=$1; shift
mkdir -p variants
BBD_LIB_OUT=variants/$name.so BBD_NVCC_EXTRA="$*" python -c "from baseboostdepth_b200 import build; build.build(force=True, verbose=True)" 2>&1 |
  grep -A2 -E "Compiling entry function '_ZN3bbd(20reproj_stream_kernelILi2ELb1|13reproj_kernelILb1ELb1)" | grep -E "Used|spill" | tr '\n' ' '
echo " <- $name"
