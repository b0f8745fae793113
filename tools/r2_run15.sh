#!/bin/bash
# round 2, GPU call 15: per-triangle shading record vs the index gathers, shade block size, at the new default batch size
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r2m_ab.log
run() { # lib, label, env...
  lib=$1; label=$2; shift 2
  echo "=== $label" >> $O/r2m_ab.log
  for w in "c5 16" "c2 0" "c3 32"; do
    set -- $w "$@"; wl=$1; spp=$2; shift 2
    env PTRS_B200_LIB=$PWD/$lib "$@" timeout 600 python bench.py --workload $wl --spp $spp --steps 3 --warmup 2 --no-cpu-baseline --no-bvh-microbench 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$wl value %.1fM e2e %.1fM ms %.1f'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step']), {k[3:]:round(v,1) for k,v in d['stage_ms'].items()})" >> $O/r2m_ab.log 2>&1
  done
}
CUR=pathtracer_rs_b200/lib/libptrs_b200.so
run build/variants/libptrs_b200_head13.so head_b24
run $CUR pack_b27
run build/variants/libptrs_b200_nopack.so nopack_b27
run build/variants/libptrs_b200_blk64.so pack_blk64_b27
run $CUR pack_b24 PTRS_PATHS_PER_BATCH=16777216
run build/variants/libptrs_b200_nopack.so nopack_b24 PTRS_PATHS_PER_BATCH=16777216
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2m_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2m_pytest.log; tail -n 4 $O/r2m_pytest.log
cat $O/r2m_ab.log
