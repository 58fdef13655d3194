#!/bin/bash
# Build an experimental variant of libcgb200.so with extra -D flags: tools/build_variant.sh NAME -DFOO=1 ...
# Output: build/variants/libcgb200_NAME.so (select with CGB_LIB=...).  Developer A/B timing only.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=build/variants/$name; mkdir -p $out
for f in cgb_sweep cgb_wsweep cgb_ewald cgb_van cgb_api; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c coulombgas_b200/csrc/$f.cu -o $out/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared --cudart static -o build/variants/libcgb200_$name.so $out/*.o
echo built build/variants/libcgb200_$name.so
