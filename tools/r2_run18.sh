#!/bin/bash
# round 2, GPU call 18: device BVH by PLOC clustering vs the radix tree (tests, build time, traversal quality against the host SAH tree)
set -u
O=gpurun_out
mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "device_bvh or tiny or c4 or 10m or ten_million" ) > $O/r2r_pytest_bvh.log 2>&1
echo "pytest rc=$?" >> $O/r2r_pytest_bvh.log; tail -n 15 $O/r2r_pytest_bvh.log
rm -f $O/r2r_bvh.log
echo "=== host SAH tree, terrain 10M" >> $O/r2r_bvh.log
timeout 300 python tools/microbench.py --all --iters 5 >> $O/r2r_bvh.log 2>&1
for b in ploc lbvh; do
  echo "=== device $b, terrain 10M" >> $O/r2r_bvh.log
  PTRS_BVH_BUILDER=$b timeout 300 python tools/microbench.py --all --iters 5 --device-bvh >> $O/r2r_bvh.log 2>&1
done
for sc in "4 262144" "2 1000000"; do
  set -- $sc
  echo "=== host SAH tree, scene $1 tris $2" >> $O/r2r_bvh.log
  timeout 300 python tools/microbench.py --all --iters 5 --scene $1 --tris $2 --side 2048 >> $O/r2r_bvh.log 2>&1
  for b in ploc lbvh; do
    echo "=== device $b, scene $1 tris $2" >> $O/r2r_bvh.log
    PTRS_BVH_BUILDER=$b timeout 300 python tools/microbench.py --all --iters 5 --scene $1 --tris $2 --side 2048 --device-bvh >> $O/r2r_bvh.log 2>&1
  done
done
cat $O/r2r_bvh.log
