#!/bin/bash
# round 2, GPU call 8: whole GPU suite at HEAD; device BVH build per kernel (launch list with DRAM bytes); C2 line with device-built tables
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r2g_parity_report.jsonl
( time PTRS_PARITY_REPORT=$PWD/$O/r2g_parity_report.jsonl timeout 1500 python -m pytest tests -m gpu -q ) > $O/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2g_pytest.log; tail -n 8 $O/r2g_pytest.log
CMD3="python tools/bvh_build_time.py --reps 3"
$CMD3 > $O/r2g_bvh_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file $O/r2g_bvh_launches.csv python tools/bvh_build_time.py --reps 1 > $O/r2g_ncu_bvh.log 2>&1
cat $O/r2g_bvh_plain.log
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-bvh-microbench > $O/bench_r2_c2.json 2> $O/bench_r2_c2.err
python - <<P
import json
d=json.load(open("$O/bench_r2_c2.json"))
print("c2 value %.1fM e2e %.1fM ms %.1f"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"]), d["e2e"]["parts_last_step"], d["e2e"]["h2d_bytes_per_step"], d["cpu_baseline"])
P
