#!/bin/bash
# Development aid: build kernel variants (compile-time knobs) side by side for one-call GPU sweeps.
#   scripts/build_variants.sh [-f cx_generic_kernels] "-DKNOB=1" "-DKNOB=2 -DOTHER=3" ...
# writes campx_b200/lib/variants/v<i>.so; select one with CAMPX_B200_LIB=<path> (campx_b200/_native.py).
set -e
cd "$(dirname "$0")/../campx_b200/csrc"
FILE=cx_agent_kernels
if [ "$1" = "-f" ]; then FILE=$2; shift 2; fi
mkdir -p ../lib/variants build/var
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC,-fvisibility=hidden -I../../include --expt-relaxed-constexpr"
OBJS=""
for f in cx_game cx_agent_kernels cx_agent_obs_kernels cx_agent_lane_kernels cx_agent_policy_kernels cx_agent_step_kernels cx_generic_kernels cx_aux_kernels cx_board_mapper; do
  [ "$f" = "$FILE" ] || OBJS="$OBJS build/$f.o"
done
i=0
for knobs in "$@"; do
  i=$((i+1))
  nvcc $FLAGS $knobs -c $FILE.cu -o build/var/${FILE}_$i.o
  nvcc -shared $ARCH -o ../lib/variants/v$i.so $OBJS build/var/${FILE}_$i.o -Xcompiler -fPIC -cudart static
  echo "v$i: $knobs"
done
