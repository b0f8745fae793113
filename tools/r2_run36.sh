#!/bin/bash
# round 2, GPU call 36: node-state tests through the PT_NO_NODE encoding, redundant PT_RB_LIVE tests dropped
set -u
O=gpurun_out
mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2ag_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2ag_pytest.log; tail -n 3 $O/r2ag_pytest.log
rm -f $O/r2ag.log
timeout 300 python tools/microbench.py --all --iters 7 2>&1 | awk '{print $1,$2,$5,$6,$7,$8}' | tr '\n' ';' >> $O/r2ag.log; echo >> $O/r2ag.log
timeout 300 python tools/render_once.py --scene 4 --tris 262144 --res 3840 2160 --spp 8 --reps 2 >> $O/r2ag.log 2>&1
timeout 300 python tools/render_once.py --scene 2 --tris 1000000 --res 1920 1080 --spp 16 --reps 2 >> $O/r2ag.log 2>&1
timeout 300 python tools/render_once.py --scene 1 --res 1024 1024 --spp 16 --reps 2 >> $O/r2ag.log 2>&1
cat $O/r2ag.log
