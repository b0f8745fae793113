#!/bin/bash
# round 2, GPU call 32: full GPU test suite, smoke, bench lines at HEAD (default with CPU baseline + microbench, C1 / C2 / C3, whole run over the device-built tree)
set -u
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r2ad_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2ad_pytest.log; tail -n 4 $O/r2ad_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -2
( time timeout 900 python bench.py ) > $O/bench_r2_final2.json 2> $O/bench_r2_final2.err; echo "bench rc=$?"; tail -3 $O/bench_r2_final2.err
for w in c1 c2 c3; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-bvh-microbench > $O/bench_r2_final2_$w.json 2>/dev/null; echo "$w rc=$?"
done
timeout 600 python bench.py --tree device --steps 2 --warmup 2 --no-cpu-baseline --no-bvh-microbench > $O/bench_r2_final2_device_tree.json 2>/dev/null; echo "device tree rc=$?"
python - <<PY
import json
for f in ("bench_r2_final2","bench_r2_final2_c1","bench_r2_final2_c2","bench_r2_final2_c3","bench_r2_final2_device_tree"):
    try:
        d=json.loads(open("$O/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.2fM"%(d["value"]/1e6), "e2e %.2fM"%(d["e2e"]["value"]/1e6), "ms %.1f"%d.get("ms_per_step",0), d.get("clocks"), {k:round(v,1) for k,v in d.get("stage_ms",{}).items()})
    except Exception as e:
        print(f, "no line", e)
PY
