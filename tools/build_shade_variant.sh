#!/bin/bash
# tools/build_shade_variant.sh <tag> <extra nvcc flags...>: rebuilds the six default shade kernels (k_shade.cu, one unit
# per material) with the flags and links a tuning variant build/variants/libptrs_b200_<tag>.so (PTRS_B200_LIB=...).
set -e
tag=$1; shift
cd "$(dirname "$0")/.."
D=build/variants/obj_$tag
mkdir -p $D
NV="/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -ftz=false -ccbin /usr/bin/g++ -Xcompiler -fPIC -diag-suppress 177 -Xptxas -v -fmad=true -prec-div=false -prec-sqrt=false"
for m in 0 1 2 3 4 5; do
  $NV "$@" -DPT_SHADE_MAT=$m -c pathtracer_rs_b200/csrc/k_shade.cu -o $D/k_shade_$m.o > $D/k_shade_$m.log 2>&1 &
done
wait
objs=$(ls build/obj/*.o | grep -v -E "k_shade_[0-5].o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o build/variants/libptrs_b200_$tag.so $objs $D/k_shade_[0-5].o -cudart static -ldl -lpthread
grep -h -A1 "Compiling entry function '_ZN4ptrs12shade_kernel" $D/k_shade_*.log | grep "stack frame"
