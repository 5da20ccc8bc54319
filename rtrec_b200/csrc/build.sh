#!/bin/sh
# Builds librtrec_b200.so for sm_100a (nvcc cross-compiles without a GPU).
# Usage: rtrec_b200/csrc/build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
mkdir -p _obj
pids=""
for f in api gram gram2 gram3 split solve wmat score score2 score3 score_tc gram_tc store upload eval; do
  if [ ! -f _obj/$f.o ] || [ $f.cu -nt _obj/$f.o ] || [ common.cuh -nt _obj/$f.o ] || [ tc_common.cuh -nt _obj/$f.o ] || [ block_select.cuh -nt _obj/$f.o ] || [ ../../include/rtrec_b200.h -nt _obj/$f.o ]; then
    $NVCC $FLAGS "$@" -c $f.cu -o _obj/$f.o &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o librtrec_b200.so _obj/api.o _obj/gram.o _obj/gram2.o _obj/gram3.o _obj/split.o _obj/solve.o _obj/wmat.o _obj/score.o _obj/score2.o _obj/score3.o _obj/score_tc.o _obj/gram_tc.o _obj/store.o _obj/upload.o _obj/eval.o
echo "built $(pwd)/librtrec_b200.so"
