#!/bin/bash
# A/B of the in-tree library against every build/variants/libptrs_b200_*.so, interleaved twice (box-to-box and
# run-to-run clock differences are larger than many of the effects being measured)
for rep in 1 2; do
for lib in pathtracer_rs_b200/lib/libptrs_b200.so build/variants/libptrs_b200_*.so; do
  echo "=== $lib"
  PTRS_B200_LIB=$PWD/$lib python tools/render_once.py --scene 1 --res 1024 1024 --spp 16 --reps 3 | cut -c40-
  PTRS_B200_LIB=$PWD/$lib python tools/render_once.py --scene 4 --res 1920 1080 --spp 4 --reps 3 | cut -c40-
  PTRS_B200_LIB=$PWD/$lib python tools/render_once.py --scene 2 --res 1920 1080 --spp 4 --reps 3 | cut -c40-
done
done
