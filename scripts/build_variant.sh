#!/bin/bash
# scripts/build_variant.sh NAME "EXTRA NVCC FLAGS": a second build of librfwb200 with compile-time knobs changed, for A/B runs on
# the GPU box (RFWB200_LIB=build_variants/lib_NAME.so python scripts/...).  Objects and the library go to build_variants/
# (git-ignored through *.so / *.o, shipped to the box by gpurun).
set -e
NAME=$1; EXTRA=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/rfw_rs_b200/csrc
OUT=$ROOT/build_variants
mkdir -p $OUT/obj_$NAME
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -ccbin /usr/bin/g++ --expt-relaxed-constexpr"
pids=()
for f in api backend builder comm radix_sort trace wavefront; do
  $NVCC $FLAGS $EXTRA -c $SRC/$f.cu -o $OUT/obj_$NAME/$f.o 2> $OUT/obj_$NAME/$f.log &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC $ARCH -shared -o $OUT/lib_$NAME.so $OUT/obj_$NAME/*.o -ldl
echo "built $OUT/lib_$NAME.so"
