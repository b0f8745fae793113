#!/bin/bash
# round 2, GPU call 19: PLOC round trace on the 10 M-triangle terrain, device-BVH tests
set -u
O=gpurun_out
mkdir -p $O
( PTRS_BVH_BUILDER=ploc PTRS_BVH_DEBUG=1 timeout 300 python tools/bvh_build_time.py --reps 2 ) > $O/r2s_ploc_trace.log 2>&1
grep -c "ploc round" $O/r2s_ploc_trace.log; grep -v "ploc round" $O/r2s_ploc_trace.log; grep "ploc round" $O/r2s_ploc_trace.log | awk 'NR<=60 || NR%10==0' | head -120
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "device_bvh or tiny" ) > $O/r2s_pytest_bvh.log 2>&1
echo "pytest rc=$?" >> $O/r2s_pytest_bvh.log; tail -n 6 $O/r2s_pytest_bvh.log
