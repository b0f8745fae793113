#!/bin/bash
# full captures of the extend and connect kernels on the C5 atrium (4K, 4 spp): launch 2 = second bounce (incoherent rays)
set -u
TAG=${1:-r1s4}
O=gpurun_out
CMD="python tools/render_once.py --scene 4 --res 3840 2160 --spp 2 --tris 262144 --reps 1"
$CMD > $O/plain_c5_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:extend_kernelILb0 -s 1 -c 1 -f -o $O/prof_c5_extend_$TAG $CMD > $O/ncu_c5_extend_$TAG.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:connect_kernelILb0 -s 1 -c 1 -f -o $O/prof_c5_connect_$TAG $CMD > $O/ncu_c5_connect_$TAG.log 2>&1
cat $O/plain_c5_$TAG.log; for f in $O/ncu_c5_*_$TAG.log; do tail -n 1 $f; done
