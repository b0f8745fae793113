#!/bin/bash
# round 2, GPU call 33: ncu --set full captures of the traversal kernels with the folded slab test (C5 extend / connect, second-bounce launch; 10 M incoherent closest)
set -u
O=gpurun_out
mkdir -p $O
TAG=r2f
CMD="python tools/render_once.py --scene 4 --res 3840 2160 --spp 2 --tris 262144 --reps 1"
timeout 300 $CMD > $O/plain_c5_$TAG.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:extend_kernelILb0ELb0 -s 1 -c 1 -f -o $O/prof_c5_extend_$TAG $CMD > $O/ncu_c5_extend_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:connect_kernelILb0ELb0 -s 1 -c 1 -f -o $O/prof_c5_connect_$TAG $CMD > $O/ncu_c5_connect_$TAG.log 2>&1
CMD2="python tools/microbench.py --rays incoherent --iters 2"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:intersect_kernelILb0ELb0ELb0 -s 1 -c 1 -f -o $O/prof_intersect_$TAG $CMD2 > $O/ncu_intersect_$TAG.log 2>&1
cat $O/plain_c5_$TAG.log; for f in $O/ncu_c5_*_$TAG.log $O/ncu_intersect_$TAG.log; do tail -n 1 $f; done
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv,noheader
ls -la $O/*.ncu-rep
