#!/bin/bash
# runs the microbench + a C2 / C5 render for every tuning variant in build/variants (PTRS_B200_LIB selects the library)
for lib in pathtracer_rs_b200/lib/libptrs_b200.so build/variants/libptrs_b200_*.so; do
  echo "=== $lib"
  PTRS_B200_LIB=$PWD/$lib python tools/microbench.py --all --iters 4 | awk '{print $1,$2,$5,$6}' | tr '\n' ';'; echo
  PTRS_B200_LIB=$PWD/$lib python tools/render_once.py --scene 1 --res 1024 1024 --spp 16 --reps 2 | cut -c40-
  PTRS_B200_LIB=$PWD/$lib python tools/render_once.py --scene 4 --res 1920 1080 --spp 4 --reps 2 | cut -c40-
done
