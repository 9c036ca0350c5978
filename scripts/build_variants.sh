#!/bin/bash
# Development aid: build kernel variants (compile-time knobs) side by side for one-call GPU sweeps.
set -e
cd "$(dirname "$0")/../campx_b200/csrc"
mkdir -p ../lib/variants build/var
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC,-fvisibility=hidden -I../../include --expt-relaxed-constexpr"
i=0
for knobs in "$@"; do
  i=$((i+1))
  nvcc $FLAGS $knobs -c cx_agent_kernels.cu -o build/var/agent_$i.o
  nvcc -shared $ARCH -o ../lib/variants/v$i.so build/cx_game.o build/var/agent_$i.o build/cx_agent_obs_kernels.o build/cx_generic_kernels.o build/cx_aux_kernels.o -Xcompiler -fPIC -cudart static
  echo "v$i: $knobs"
done
