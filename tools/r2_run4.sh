#!/bin/bash
# round 2, GPU call 4: probe / parity / device-table tests with the parity report; compute-sanitizer memcheck of the Cornell + env render
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r2d_parity_report.jsonl
( PTRS_PARITY_REPORT=$PWD/$O/r2d_parity_report.jsonl timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -m gpu -q -k "bxdf or lights or exact_shading or device_built or multi or comm or deeper" ) > $O/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2d_pytest.log; tail -n 25 $O/r2d_pytest.log
CMD="examples/headless tests/golden/cornell-box-sunsky.xml -o /tmp/san_out --headless -r 160x120 -s 4 -d 8 --server 127.0.0.1:1 --sunsky-hdr tests/golden/abandoned_tank_farm_04_1k.hdr"
mkdir -p /tmp/san_out
$CMD > $O/r2d_headless_plain.log 2>&1 && timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 $CMD > $O/r2d_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -n 6 $O/r2d_memcheck.log
