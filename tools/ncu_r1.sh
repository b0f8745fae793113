#!/bin/bash
# Round-1 profiling pass (run under gpurun): launch list of one bench step, then full captures of the
# extend kernel (render) and the standalone closest-hit kernel on incoherent rays (10 M triangles).
set -u
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-bvh-microbench"
$CMD > gpurun_out/plain_render.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain_render2.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:extend_kernelILb0 -s 1 -c 3 -f -o gpurun_out/prof_extend_r1 $CMD > gpurun_out/ncu_extend.log 2>&1
CMD2="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
$CMD2 > gpurun_out/plain_micro.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:intersect_kernelILb0ELb0 -s 8 -c 2 -f -o gpurun_out/prof_intersect_incoherent_r1 $CMD2 > gpurun_out/ncu_intersect.log 2>&1
tail -3 gpurun_out/ncu_list.log gpurun_out/ncu_extend.log gpurun_out/ncu_intersect.log
