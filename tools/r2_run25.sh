#!/bin/bash
# round 2, GPU call 25: traversal kernels at 10 / 12 blocks per SM (48 / 40 registers, spilling) against 8 (64 registers)
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r2y_ab.log
for lib in pathtracer_rs_b200/lib/libptrs_b200.so build/variants/libptrs_b200_mb10.so build/variants/libptrs_b200_mb12.so; do
  echo "=== $lib" >> $O/r2y_ab.log
  PTRS_B200_LIB=$PWD/$lib timeout 300 python tools/microbench.py --all --iters 5 2>&1 | awk '{print $1,$2,$5,$6,$7,$8}' | tr '\n' ';' >> $O/r2y_ab.log; echo >> $O/r2y_ab.log
  PTRS_B200_LIB=$PWD/$lib timeout 300 python tools/render_once.py --scene 4 --tris 262144 --res 3840 2160 --spp 8 --reps 2 >> $O/r2y_ab.log 2>&1
done
cat $O/r2y_ab.log
