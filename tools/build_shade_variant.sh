#!/bin/bash
# tools/build_shade_variant.sh <tag> <extra nvcc flags...>: rebuilds the six shade units with the flags and links
# a tuning variant build/variants/libptrs_b200_<tag>.so (select it with PTRS_B200_LIB=...).
set -e
tag=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/variants/obj_$tag
make -s -j8 OBJ=build/variants/obj_$tag LIB=build/variants/lib_$tag SHADE_EXTRA="$*" build/variants/lib_$tag/libptrs_b200.so > /dev/null
cp build/variants/lib_$tag/libptrs_b200.so build/variants/libptrs_b200_$tag.so
for m in 0 1 2 3 4 5; do grep -A2 "shade_kernel" build/variants/obj_$tag/k_shade_$m.ptxas.log | grep -E "Used" | awk '{print $5}' | tr '\n' ' '; done; echo
