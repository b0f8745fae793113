#!/bin/bash
# round 2, GPU call 40: near-child bit taken with one shift less; full GPU suite + timings
set -u
O=gpurun_out
mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2ai_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2ai_pytest.log; tail -n 3 $O/r2ai_pytest.log
rm -f $O/r2ai.log
timeout 300 python tools/microbench.py --all --iters 7 2>&1 | awk '{print $1,$2,$5,$6,$7,$8}' | tr '\n' ';' >> $O/r2ai.log; echo >> $O/r2ai.log
timeout 300 python tools/render_once.py --scene 4 --tris 262144 --res 3840 2160 --spp 8 --reps 2 >> $O/r2ai.log 2>&1
cat $O/r2ai.log
