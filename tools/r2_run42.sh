#!/bin/bash
# round 2, GPU call 42: node records loaded with the L1 evict_last hint
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r2ak.log
for lib in pathtracer_rs_b200/lib/libptrs_b200.so build/variants/libptrs_b200_evl.so; do
  echo "=== $lib" >> $O/r2ak.log
  PTRS_B200_LIB=$PWD/$lib timeout 200 python tools/microbench.py --all --iters 5 2>&1 | awk '{print $1,$2,$5,$6,$7,$8}' | tr '\n' ';' >> $O/r2ak.log; echo >> $O/r2ak.log
  PTRS_B200_LIB=$PWD/$lib timeout 200 python tools/render_once.py --scene 4 --tris 262144 --res 3840 2160 --spp 8 --reps 2 >> $O/r2ak.log 2>&1
done
cat $O/r2ak.log
