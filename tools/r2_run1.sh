#!/bin/bash
# round 2, GPU call 1: full GPU test suite, traversal A/B (round-1 engine, 16-byte stack, current), default bench, launch list
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2a_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2a_pytest.log
tail -n 5 $O/r2a_pytest.log
for lib in pathtracer_rs_b200/lib/libptrs_b200.so build/variants/libptrs_b200_r1trace.so build/variants/libptrs_b200_stack16.so; do
  echo "=== $lib" >> $O/r2a_ab.log
  PTRS_B200_LIB=$PWD/$lib timeout 600 python tools/microbench.py --all --iters 4 >> $O/r2a_ab.log 2>&1
  PTRS_B200_LIB=$PWD/$lib timeout 600 python bench.py --workload c5 --spp 16 --steps 3 --warmup 2 --no-cpu-baseline --no-bvh-microbench 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c5@16spp value %.1fM ms %.1f'%(d['value']/1e6,d['ms_per_step']), {k[3:]:round(v,1) for k,v in d['stage_ms'].items()}, 'trav frac', round(d['roofline']['frac'],3), round(d['roofline'].get('frac_of_l2_gather',0),3))" >> $O/r2a_ab.log 2>&1
  PTRS_B200_LIB=$PWD/$lib timeout 600 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu-baseline --no-bvh-microbench 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c2 value %.1fM ms %.1f'%(d['value']/1e6,d['ms_per_step']), {k[3:]:round(v,1) for k,v in d['stage_ms'].items()})" >> $O/r2a_ab.log 2>&1
done
cat $O/r2a_ab.log
( time timeout 1200 python bench.py --steps 3 --warmup 1 ) > $O/r2a_bench.json 2> $O/r2a_bench.err
echo "bench rc=$?"; head -c 600 $O/r2a_bench.json
CMD="python bench.py --workload c5 --spp 8 --steps 1 --warmup 0 --no-cpu-baseline --no-bvh-microbench"
$CMD > $O/r2a_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r2a_launches_c5.csv $CMD > $O/r2a_ncu_launch.log 2>&1
echo "ncu rc=$?"
