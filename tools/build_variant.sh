#!/bin/bash
# Builds an experimental variant of the library: tools/build_variant.sh <name> [-D...]  -> gpurun_variants/lib_<name>.so
# (select it at run time with MODA_B200_LIB=gpurun_variants/lib_<name>.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p gpurun_variants/obj_$name
for f in api elementwise skin composite gemm sample_pdf raysum sinkhorn flow tc_gemm tc_support chain; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
    "$@" -I moda_b200/csrc -c moda_b200/csrc/$f.cu -o gpurun_variants/obj_$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o gpurun_variants/lib_$name.so gpurun_variants/obj_$name/*.o -lcudart -lcuda
rm -rf gpurun_variants/obj_$name
echo gpurun_variants/lib_$name.so
