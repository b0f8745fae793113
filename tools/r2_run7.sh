#!/bin/bash
# round 2, GPU call 7: whole GPU suite at HEAD (infinite-light MIS rays as any-hit rays, SAH-guided leaf collapse of the device BVH),
# stage times of C5 / C2 / C3, device-built BVH on the 10 M-triangle mesh
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r2f_parity_report.jsonl
( time PTRS_PARITY_REPORT=$PWD/$O/r2f_parity_report.jsonl timeout 1500 python -m pytest tests -m gpu -q ) > $O/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2f_pytest.log; tail -n 8 $O/r2f_pytest.log
rm -f $O/r2f_stage.log
for w in "c5 16" "c2 0" "c3 32" "c1 0"; do
  set -- $w; wl=$1; spp=$2
  timeout 600 python bench.py --workload $wl --spp $spp --steps 3 --warmup 2 --no-cpu-baseline --no-bvh-microbench 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$wl value %.1fM e2e %.1fM ms %.1f'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step']), {k[3:]:round(v,1) for k,v in d['stage_ms'].items()}, d['rays_rank0'], 'frac', round(d['roofline']['frac'],3), round(d['roofline'].get('frac_of_l2_gather',0),3))" >> $O/r2f_stage.log 2>&1
done
timeout 600 python tools/microbench.py --rays incoherent --iters 4 --device-bvh >> $O/r2f_stage.log 2>&1
timeout 600 python tools/microbench.py --rays coherent --iters 4 --device-bvh 2>&1 | tail -n 1 >> $O/r2f_stage.log
cat $O/r2f_stage.log
