#!/bin/bash
# round 2, GPU call 2: ray reordering between bounces on / off and key variants
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_analytic_anchors.py -m gpu -x -q -k "not full_size and not c4_hit" ) > $O/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2b_pytest.log; tail -n 4 $O/r2b_pytest.log
run() {  # label, env...
  label=$1; shift
  for w in "c5 16" "c2 0" "c3 32"; do
    set -- $w "$@"
    wl=$1; spp=$2; shift 2
    env "$@" timeout 600 python bench.py --workload $wl --spp $spp --steps 3 --warmup 2 --no-cpu-baseline --no-bvh-microbench 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$label $wl value %.1fM ms %.1f'%(d['value']/1e6,d['ms_per_step']), {k[3:]:round(v,1) for k,v in d['stage_ms'].items()}, 'launches', d['gpu_launches']//3)" >> $O/r2b_sort.log 2>&1
  done
}
run off PTRS_SORT_RAYS=0
run on PTRS_SORT_RAYS=1
run on_bit5 PTRS_SORT_RAYS=1 PTRS_SORT_BEGIN_BIT=5
run on_bit14 PTRS_SORT_RAYS=1 PTRS_SORT_BEGIN_BIT=14
run on_min256k PTRS_SORT_RAYS=1 PTRS_SORT_MIN=262144
cat $O/r2b_sort.log
