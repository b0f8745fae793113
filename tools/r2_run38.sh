#!/bin/bash
# round 2, GPU call 38: first triangle of a parked leaf prefetched into L1 at park time
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r2ah.log
for lib in pathtracer_rs_b200/lib/libptrs_b200.so build/variants/libptrs_b200_parkpf.so; do
  echo "=== $lib" >> $O/r2ah.log
  PTRS_B200_LIB=$PWD/$lib timeout 300 python tools/microbench.py --all --iters 5 2>&1 | awk '{print $1,$2,$5,$6,$7,$8}' | tr '\n' ';' >> $O/r2ah.log; echo >> $O/r2ah.log
  PTRS_B200_LIB=$PWD/$lib timeout 300 python tools/render_once.py --scene 4 --tris 262144 --res 3840 2160 --spp 8 --reps 2 >> $O/r2ah.log 2>&1
  PTRS_B200_LIB=$PWD/$lib timeout 300 python tools/render_once.py --scene 2 --tris 1000000 --res 1920 1080 --spp 16 --reps 2 >> $O/r2ah.log 2>&1
done
cat $O/r2ah.log
