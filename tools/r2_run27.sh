#!/bin/bash
# round 2, GPU call 27: 8-byte stack entries and 4 box steps per decision at HEAD (candidates for trees beyond L2 size)
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r2z_ab.log
for lib in pathtracer_rs_b200/lib/libptrs_b200.so build/variants/libptrs_b200_stack8.so build/variants/libptrs_b200_stack8box4.so build/variants/libptrs_b200_box4.so; do
  echo "=== $lib" >> $O/r2z_ab.log
  PTRS_B200_LIB=$PWD/$lib timeout 300 python tools/microbench.py --all --iters 7 2>&1 | awk '{print $1,$2,$5,$6,$7,$8}' | tr '\n' ';' >> $O/r2z_ab.log; echo >> $O/r2z_ab.log
  PTRS_B200_LIB=$PWD/$lib timeout 300 python tools/render_once.py --scene 4 --tris 262144 --res 3840 2160 --spp 8 --reps 2 >> $O/r2z_ab.log 2>&1
  PTRS_B200_LIB=$PWD/$lib timeout 300 python tools/render_once.py --scene 3 --tris 10000000 --res 1920 1080 --spp 16 --reps 2 >> $O/r2z_ab.log 2>&1
done
( PTRS_B200_LIB=$PWD/build/variants/libptrs_b200_stack8box4.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "intersect or traversal or deterministic or c4" ) > $O/r2z_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2z_pytest.log; tail -n 3 $O/r2z_pytest.log
cat $O/r2z_ab.log
