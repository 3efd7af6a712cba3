#!/bin/bash
# Build A/B variants of libbooster_b200.so from the same sources (HERE, before gpurun: nvcc cross-compiles):
#   bash scripts/gpu_variants.sh name "-DB200_X=1 ..." [name "-D.."]...   -> build/variants/libbooster_b200_<name>.so
# On the GPU box select one with BOOSTER_B200_LIB=build/variants/libbooster_b200_<name>.so (diagnostic only).
set -e
mkdir -p build/variants
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function $defs \
       -c booster_b200/csrc/engine.cu -o build/variants/engine_$name.o 2> build/variants/engine_$name.log
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libbooster_b200_$name.so build/variants/engine_$name.o build/gguf.o build/bridge.o build/tokenizer.o -ldl
  echo "built $name ($defs)"
done
