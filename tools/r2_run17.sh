#!/bin/bash
# round 2, GPU call 17: bottom K entries of the traversal stack in shared memory (PT_SMEM_STACK = 4 / 8 / 12), with 3 and 4 box steps
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r2q_ab.log
for lib in pathtracer_rs_b200/lib/libptrs_b200.so build/variants/libptrs_b200_smem4.so build/variants/libptrs_b200_smem8.so build/variants/libptrs_b200_smem12.so build/variants/libptrs_b200_smem8box4.so; do
  echo "=== $lib" >> $O/r2q_ab.log
  PTRS_B200_LIB=$PWD/$lib timeout 300 python tools/microbench.py --all --iters 5 2>&1 | awk '{print $1,$2,$5,$6,$7,$8}' | tr '\n' ';' >> $O/r2q_ab.log; echo >> $O/r2q_ab.log
  for w in "c5 16" "c2 0"; do
    set -- $w; wl=$1; spp=$2
    env PTRS_B200_LIB=$PWD/$lib timeout 600 python bench.py --workload $wl --spp $spp --steps 3 --warmup 2 --no-cpu-baseline --no-bvh-microbench 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$wl value %.1fM e2e %.1fM ms %.1f'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step']), {k[3:]:round(v,1) for k,v in d['stage_ms'].items()})" >> $O/r2q_ab.log 2>&1
  done
done
( PTRS_B200_LIB=$PWD/build/variants/libptrs_b200_smem8.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "intersect or traversal or deterministic" ) > $O/r2q_pytest_smem8.log 2>&1
echo "pytest rc=$?" >> $O/r2q_pytest_smem8.log; tail -n 4 $O/r2q_pytest_smem8.log
cat $O/r2q_ab.log
