#!/bin/bash
# round 2, GPU call 34: scheduling knobs re-swept on the folded box step (box steps per decision 2 / 3 / 4, box_min 16 / 20 / 24)
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r2ae.log
run() { # lib, label, env...
  lib=$1; label=$2; shift 2
  echo "=== $label" >> $O/r2ae.log
  env PTRS_B200_LIB=$PWD/$lib "$@" timeout 300 python tools/microbench.py --all --iters 5 2>&1 | awk '{print $1,$2,$5,$6,$7,$8}' | tr '\n' ';' >> $O/r2ae.log; echo >> $O/r2ae.log
  env PTRS_B200_LIB=$PWD/$lib "$@" timeout 300 python tools/render_once.py --scene 4 --tris 262144 --res 3840 2160 --spp 8 --reps 2 >> $O/r2ae.log 2>&1
}
CUR=pathtracer_rs_b200/lib/libptrs_b200.so
run $CUR box3_min20
run $CUR box3_min16 PTRS_BOX_MIN=16
run $CUR box3_min24 PTRS_BOX_MIN=24
run build/variants/libptrs_b200_box4.so box4_min20
run build/variants/libptrs_b200_box2.so box2_min20
run build/variants/libptrs_b200_box4.so box4_min24 PTRS_BOX_MIN=24
cat $O/r2ae.log
