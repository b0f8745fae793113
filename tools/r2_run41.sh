#!/bin/bash
# round 2, GPU call 41: any-hit-only specialisation of the traversal (connect without area lights, ptrs_intersect_p): full GPU suite + timings
set -u
O=gpurun_out
mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2aj_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2aj_pytest.log; tail -n 3 $O/r2aj_pytest.log
rm -f $O/r2aj.log
timeout 300 python tools/microbench.py --all --iters 7 2>&1 | awk '{print $1,$2,$5,$6,$7,$8}' | tr '\n' ';' >> $O/r2aj.log; echo >> $O/r2aj.log
timeout 300 python tools/render_once.py --scene 4 --tris 262144 --res 3840 2160 --spp 8 --reps 2 >> $O/r2aj.log 2>&1
timeout 300 python tools/render_once.py --scene 4 --tris 262144 --res 3840 2160 --spp 8 --reps 2 --device-bvh >> $O/r2aj.log 2>&1
cat $O/r2aj.log
