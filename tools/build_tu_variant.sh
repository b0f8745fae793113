#!/bin/bash
# tools/build_tu_variant.sh <unit> <tag> <extra nvcc flags...>: rebuilds csrc/<unit>.cu (an exact-arithmetic unit: k_bvh, k_trace,
# k_misc, k_tables) with the flags and links build/variants/libptrs_b200_<tag>.so (select it with PTRS_B200_LIB=...)
set -e
unit=$1; tag=$2; shift 2
cd "$(dirname "$0")/.."
mkdir -p build/variants
NV="/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -ccbin /usr/bin/g++ -Xcompiler -fPIC -diag-suppress 177 -Xptxas -v"
$NV "$@" -c pathtracer_rs_b200/csrc/$unit.cu -o build/variants/${unit}_$tag.o > build/variants/${unit}_$tag.log 2>&1
objs=$(ls build/obj/*.o | grep -v "/$unit.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o build/variants/libptrs_b200_$tag.so $objs build/variants/${unit}_$tag.o -cudart static -ldl -lpthread
echo "built build/variants/libptrs_b200_$tag.so"
