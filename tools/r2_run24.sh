#!/bin/bash
# round 2, GPU call 24: device BVH build with the cached working block: phase trace over repeated builds, BVH / multi-GPU tests
set -u
O=gpurun_out
mkdir -p $O
( PTRS_BVH_DEBUG=1 timeout 300 python tools/bvh_build_time.py --reps 4 ) 2>&1 | grep -v "ploc round" > $O/r2x_build_trace.log
cat $O/r2x_build_trace.log
( timeout 300 python tools/bvh_build_time.py --reps 4 ) 2>&1 | tail -4
( PTRS_BVH_BUILDER=lbvh timeout 300 python tools/bvh_build_time.py --reps 4 ) 2>&1 | tail -4
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -m gpu -x -q -k "bvh or tiny or multi or c4" ) > $O/r2x_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2x_pytest.log; tail -n 4 $O/r2x_pytest.log
