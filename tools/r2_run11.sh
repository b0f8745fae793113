#!/bin/bash
# round 2, GPU call 11: box steps per scheduling decision (1 .. 4), box_min retune
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "intersect or counters or c4_hit or path_radiance or render_image or device_bvh or deeper" ) > $O/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2j_pytest.log; tail -n 4 $O/r2j_pytest.log
rm -f $O/r2j_ab.log
run() { # lib, label, env...
  lib=$1; label=$2; shift 2
  echo "=== $label" >> $O/r2j_ab.log
  env PTRS_B200_LIB=$PWD/$lib "$@" timeout 600 python tools/microbench.py --all --iters 4 2>&1 | grep -v "^tris=" | awk '{print $1,$2,$5,$6,$7,$8}' >> $O/r2j_ab.log
  for w in "c5 16" "c2 0" "c3 32"; do
    set -- $w "$@"; wl=$1; spp=$2; shift 2
    env PTRS_B200_LIB=$PWD/$lib "$@" timeout 600 python bench.py --workload $wl --spp $spp --steps 3 --warmup 2 --no-cpu-baseline --no-bvh-microbench 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$wl value %.1fM ms %.1f'%(d['value']/1e6,d['ms_per_step']), {k[3:]:round(v,1) for k,v in d['stage_ms'].items()})" >> $O/r2j_ab.log 2>&1
  done
}
run build/variants/libptrs_b200_steps1.so steps1
run pathtracer_rs_b200/lib/libptrs_b200.so steps2
run build/variants/libptrs_b200_steps3.so steps3
run build/variants/libptrs_b200_steps4.so steps4
run pathtracer_rs_b200/lib/libptrs_b200.so steps2_boxmin16 PTRS_BOX_MIN=16
run pathtracer_rs_b200/lib/libptrs_b200.so steps2_boxmin24 PTRS_BOX_MIN=24
run build/variants/libptrs_b200_steps3.so steps3_boxmin16 PTRS_BOX_MIN=16
cat $O/r2j_ab.log
