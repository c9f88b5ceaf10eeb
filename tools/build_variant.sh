#!/bin/sh
# Build another copy of the library with extra nvcc flags (experiments / instrumented builds):
#   tools/build_variant.sh NAME "-DAGB_LOOP_STATS ..."  ->  aprilgrid-rs_b200/lib/variants/libag_NAME.so
# Select it with AG_LIB=aprilgrid-rs_b200/lib/variants/libag_NAME.so (profiling tools only).
set -e
cd "$(dirname "$0")/../aprilgrid-rs_b200"
name=$1; shift
mkdir -p build/v_$name lib/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -cudart static $*"
for f in ag_api ag_dense ag_sparse ag_board ag_render; do
  nvcc $FLAGS -c csrc/$f.cu -o build/v_$name/$f.o &
done
wait
nvcc $ARCH -shared -cudart static -o lib/variants/libag_$name.so build/v_$name/*.o
echo built lib/variants/libag_$name.so
