#!/bin/bash
# Profiling pass (run under gpurun): launch list of one Cornell+env render, then full captures of the
# extend / shade<Matte> / connect kernels of that render and of the standalone closest-hit kernel on
# incoherent rays over the 10 M-triangle terrain.  usage: tools/ncu_r1b.sh <tag>
set -u
TAG=${1:-r1b}
O=gpurun_out
CMD="python tools/render_once.py --scene 1 --res 1024 1024 --spp 16 --reps 1"
$CMD > $O/plain_render_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_list_$TAG.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:extend_kernelILb0 -s 2 -c 1 -f -o $O/prof_extend_$TAG $CMD > $O/ncu_extend_$TAG.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:shade_kernelILi0 -s 2 -c 1 -f -o $O/prof_shade0_$TAG $CMD > $O/ncu_shade_$TAG.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:connect_kernelILb0 -s 2 -c 1 -f -o $O/prof_connect_$TAG $CMD > $O/ncu_connect_$TAG.log 2>&1
CMD2="python tools/microbench.py --rays incoherent --iters 2"
$CMD2 > $O/plain_micro_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:intersect_kernelILb0ELb0 -s 1 -c 1 -f -o $O/prof_intersect_$TAG $CMD2 > $O/ncu_intersect_$TAG.log 2>&1
for f in $O/ncu_list_$TAG.log $O/ncu_extend_$TAG.log $O/ncu_shade_$TAG.log $O/ncu_connect_$TAG.log $O/ncu_intersect_$TAG.log; do tail -n 2 $f; done
