#!/bin/bash
# round 2, GPU call 26: bench lines at HEAD (default = C5 with CPU baseline and BVH microbench; reference arm; C1 / C2 / C3), launch list of the bench command
set -u
O=gpurun_out
mkdir -p $O
( time timeout 900 python bench.py ) > $O/bench_r2_final.json 2> $O/bench_r2_final.err; echo "bench rc=$?"; tail -3 $O/bench_r2_final.err
( time timeout 600 python bench.py --impl reference --steps 1 --warmup 0 ) > $O/bench_r2_final_reference.json 2> $O/bench_r2_final_reference.err; echo "ref rc=$?"
for w in c1 c2 c3; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-bvh-microbench > $O/bench_r2_final_$w.json 2>/dev/null; echo "$w rc=$?"
done
BENCH="python bench.py --spp 16 --steps 2 --warmup 1 --no-cpu-baseline --no-bvh-microbench"
$BENCH > $O/plain_bench_final.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r2_final_launches_bench_c5_16spp.csv $BENCH > $O/ncu_list_final.log 2>&1
tail -n 2 $O/ncu_list_final.log
python tools/launch_summary.py $O/r2_final_launches_bench_c5_16spp.csv | cut -c1-150 | head -30
python - <<PY
import json
for f in ("bench_r2_final","bench_r2_final_reference","bench_r2_final_c1","bench_r2_final_c2","bench_r2_final_c3"):
    try:
        d=json.loads(open("$O/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.2fM"%(d["value"]/1e6), "e2e %.2fM"%(d["e2e"]["value"]/1e6), "ms %.1f"%d.get("ms_per_step",0), d.get("clocks"), {k:round(v,1) for k,v in d.get("stage_ms",{}).items()})
    except Exception as e:
        print(f, "no line", e)
PY
