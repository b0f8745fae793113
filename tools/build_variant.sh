#!/bin/bash
# tools/build_variant.sh <tag> <extra nvcc flags...>: rebuilds k_trace.cu with the flags and links a
# tuning variant build/variants/libptrs_b200_<tag>.so (select it with PTRS_B200_LIB=...).
set -e
tag=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/variants
NV="/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -ccbin /usr/bin/g++ -Xcompiler -fPIC -diag-suppress 177 -Xptxas -v"
$NV "$@" -c ${KTRACE_SRC:-pathtracer_rs_b200/csrc/k_trace.cu} -o build/variants/k_trace_$tag.o > build/variants/k_trace_$tag.log 2>&1
objs=$(ls build/obj/*.o | grep -v k_trace.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o build/variants/libptrs_b200_$tag.so $objs build/variants/k_trace_$tag.o -cudart static -ldl -lpthread
grep -E "Used" build/variants/k_trace_$tag.log | awk '{print $5}' | tr '\n' ' '; grep -c "spill stores" build/variants/k_trace_$tag.log; grep "spill" build/variants/k_trace_$tag.log | grep -v " 0 bytes spill stores" | head -3
