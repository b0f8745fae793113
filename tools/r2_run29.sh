#!/bin/bash
# round 2, GPU call 29: library-built trees walked front to back (nearer child first) against dir_is_neg[axis]
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r2aa_ab.log
for d in 0 1; do
  echo "=== PTRS_DIST_ORDER=$d" >> $O/r2aa_ab.log
  for sc in "3 10000000 4096" "4 262144 2048" "2 1000000 2048"; do
    set -- $sc
    PTRS_DIST_ORDER=$d timeout 300 python tools/microbench.py --all --iters 5 --scene $1 --tris $2 --side $3 --device-bvh 2>&1 | grep -v "^tris\|^device" | awk '{print $1,$2,$6,$7,$8,$9,$10,$11}' | tr '\n' ';' >> $O/r2aa_ab.log; echo >> $O/r2aa_ab.log
  done
  PTRS_DIST_ORDER=$d timeout 300 python tools/render_once.py --scene 4 --tris 262144 --res 3840 2160 --spp 8 --reps 2 --device-bvh >> $O/r2aa_ab.log 2>&1
  PTRS_DIST_ORDER=$d timeout 300 python tools/render_once.py --scene 2 --tris 1000000 --res 1920 1080 --spp 16 --reps 2 --device-bvh >> $O/r2aa_ab.log 2>&1
  PTRS_DIST_ORDER=$d timeout 300 python tools/render_once.py --scene 1 --res 1024 1024 --spp 16 --reps 2 --device-bvh >> $O/r2aa_ab.log 2>&1
done
echo "=== reference-built tree (unchanged path)" >> $O/r2aa_ab.log
timeout 300 python tools/render_once.py --scene 4 --tris 262144 --res 3840 2160 --spp 8 --reps 2 >> $O/r2aa_ab.log 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q ) > $O/r2aa_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2aa_pytest.log; tail -n 3 $O/r2aa_pytest.log
cat $O/r2aa_ab.log
