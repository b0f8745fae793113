#!/bin/bash
# full captures of the Disney and substrate shade kernels on the C3 material field (1 M triangles, 1920x1080, 2 spp)
set -u
TAG=${1:-r1s4}
O=gpurun_out
CMD="python tools/render_once.py --scene 2 --res 1920 1080 --spp 2 --tris 1000000 --reps 1"
$CMD > $O/plain_c3_$TAG.log 2>&1 || exit 1
for k in 5 4; do
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:shade_kernelILi$k -s 0 -c 1 -f -o $O/prof_c3_shade${k}_$TAG $CMD > $O/ncu_c3_shade${k}_$TAG.log 2>&1
done
cat $O/plain_c3_$TAG.log; for f in $O/ncu_c3_shade*_$TAG.log; do tail -n 1 $f; done
