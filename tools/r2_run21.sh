#!/bin/bash
# round 2, GPU call 21: phase trace of the device BVH build (both builders, 10 M triangles)
set -u
O=gpurun_out
mkdir -p $O
for b in ploc lbvh; do
( PTRS_BVH_BUILDER=$b PTRS_BVH_DEBUG=1 timeout 300 python tools/bvh_build_time.py --reps 3 ) > $O/r2u_trace_$b.log 2>&1
grep -v "ploc round" $O/r2u_trace_$b.log
done
