#!/bin/bash
# round 2, GPU call 23: PLOC search radius (8 / 16 / 32 / 64) and leaf size limit (2 / 4 / 8): render times over the device-built tree
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r2w_ploc_variants.log
for lib in pathtracer_rs_b200/lib/libptrs_b200.so build/variants/libptrs_b200_plocr8.so build/variants/libptrs_b200_plocr32.so build/variants/libptrs_b200_plocr64.so build/variants/libptrs_b200_leaf2.so build/variants/libptrs_b200_leaf8.so; do
  echo "=== $lib" >> $O/r2w_ploc_variants.log
  PTRS_B200_LIB=$PWD/$lib timeout 300 python tools/render_once.py --scene 4 --tris 262144 --res 3840 2160 --spp 8 --reps 2 --device-bvh >> $O/r2w_ploc_variants.log 2>&1
  PTRS_B200_LIB=$PWD/$lib timeout 300 python tools/render_once.py --scene 2 --tris 1000000 --res 1920 1080 --spp 16 --reps 2 --device-bvh >> $O/r2w_ploc_variants.log 2>&1
  PTRS_B200_LIB=$PWD/$lib timeout 300 python tools/microbench.py --all --iters 3 --device-bvh 2>&1 | awk '{print $1,$2,$5,$6,$7,$8,$9,$10,$11,$12}' | tr '\n' ';' >> $O/r2w_ploc_variants.log; echo >> $O/r2w_ploc_variants.log
done
( PTRS_BVH_DEBUG=1 timeout 300 python tools/bvh_build_time.py --reps 3 ) 2>&1 | grep -v "ploc round" >> $O/r2w_ploc_variants.log
cat $O/r2w_ploc_variants.log
