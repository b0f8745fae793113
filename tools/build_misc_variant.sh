#!/bin/bash
# tools/build_misc_variant.sh <tag> <extra nvcc flags...>: rebuilds k_misc.cu with the flags and links a tuning
# variant build/variants/libptrs_b200_<tag>.so (select it with PTRS_B200_LIB=...).
set -e
tag=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/variants
NV="/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -ccbin /usr/bin/g++ -Xcompiler -fPIC -diag-suppress 177 -Xptxas -v"
$NV "$@" -c pathtracer_rs_b200/csrc/k_misc.cu -o build/variants/k_misc_$tag.o > build/variants/k_misc_$tag.log 2>&1
objs=$(ls build/obj/*.o | grep -v k_misc.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o build/variants/libptrs_b200_$tag.so $objs build/variants/k_misc_$tag.o -cudart static
grep -A2 "connect_resolve" build/variants/k_misc_$tag.log | grep -E "Used|spill" | head -3
