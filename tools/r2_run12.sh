#!/bin/bash
# round 2, GPU call 12: triangle steps per scheduling decision, refill threshold
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "intersect or counters or c4_hit or path_radiance or render_image or device_bvh or deeper" ) > $O/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2k_pytest.log; tail -n 4 $O/r2k_pytest.log
rm -f $O/r2k_ab.log
run() { # lib, label, env...
  lib=$1; label=$2; shift 2
  echo "=== $label" >> $O/r2k_ab.log
  env PTRS_B200_LIB=$PWD/$lib "$@" timeout 600 python tools/microbench.py --all --iters 4 2>&1 | grep -v "^tris=" | awk '{print $1,$2,$5,$6,$7,$8}' >> $O/r2k_ab.log
  for w in "c5 16" "c2 0" "c3 32"; do
    set -- $w "$@"; wl=$1; spp=$2; shift 2
    env PTRS_B200_LIB=$PWD/$lib "$@" timeout 600 python bench.py --workload $wl --spp $spp --steps 3 --warmup 2 --no-cpu-baseline --no-bvh-microbench 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$wl value %.1fM ms %.1f'%(d['value']/1e6,d['ms_per_step']), {k[3:]:round(v,1) for k,v in d['stage_ms'].items()})" >> $O/r2k_ab.log 2>&1
  done
}
run pathtracer_rs_b200/lib/libptrs_b200.so head_steps3
run build/variants/libptrs_b200_tri2.so tri2
run build/variants/libptrs_b200_refill4.so refill4
run build/variants/libptrs_b200_refill12.so refill12
run build/variants/libptrs_b200_refill16.so refill16
for v in tri2 refill4; do
  PTRS_B200_LIB=$PWD/build/variants/libptrs_b200_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "intersect or counters or c4_hit or path_radiance" 2>&1 | tail -n 1 >> $O/r2k_ab.log
done
cat $O/r2k_ab.log
