#!/bin/bash
# full captures of shade<Matte> (first-bounce launch and a later, scrambled-order launch) on the C2 scene, 16 spp
set -u
TAG=${1:-r1s4}
O=gpurun_out
CMD="python tools/render_once.py --scene 1 --res 1024 1024 --spp 16 --reps 1"
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:shade_kernelILi0 -s 2 -c 1 -f -o $O/prof_shade0_$TAG $CMD > $O/ncu_shade0_$TAG.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:shade_kernelILi0 -s 5 -c 1 -f -o $O/prof_shade0b3_$TAG $CMD > $O/ncu_shade0b3_$TAG.log 2>&1
for f in $O/ncu_shade0*_$TAG.log; do tail -n 1 $f; done
