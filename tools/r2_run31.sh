#!/bin/bash
# round 2, GPU call 31: folded slab test with the node pair fetched as four 128-bit loads (default) or two 256-bit loads
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r2ac.log
for lib in pathtracer_rs_b200/lib/libptrs_b200.so build/variants/libptrs_b200_ld256.so; do
  echo "=== $lib" >> $O/r2ac.log
  PTRS_B200_LIB=$PWD/$lib timeout 300 python tools/microbench.py --all --iters 7 2>&1 | awk '{print $1,$2,$5,$6,$7,$8}' | tr '\n' ';' >> $O/r2ac.log; echo >> $O/r2ac.log
  PTRS_B200_LIB=$PWD/$lib timeout 300 python tools/render_once.py --scene 4 --tris 262144 --res 3840 2160 --spp 8 --reps 2 >> $O/r2ac.log 2>&1
  PTRS_B200_LIB=$PWD/$lib timeout 300 python tools/render_once.py --scene 2 --tris 1000000 --res 1920 1080 --spp 16 --reps 2 >> $O/r2ac.log 2>&1
  PTRS_B200_LIB=$PWD/$lib timeout 300 python tools/render_once.py --scene 1 --res 1024 1024 --spp 16 --reps 2 >> $O/r2ac.log 2>&1
done
( timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "intersect or traversal or c4 or deterministic or counters" ) > $O/r2ac_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2ac_pytest.log; tail -n 3 $O/r2ac_pytest.log
cat $O/r2ac.log
