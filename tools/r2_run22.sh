#!/bin/bash
# round 2, GPU call 22: default bench line with the device-built (PLOC) tree beside the reference-built one, C3 / C2 the same way, full GPU test suite
set -u
O=gpurun_out
mkdir -p $O
( time timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-bvh-microbench ) > $O/r2v_bench_c5.json 2> $O/r2v_bench_c5.err
echo "c5 rc=$?"; python - <<PY
import json
d=json.loads(open("$O/r2v_bench_c5.json").read().strip().splitlines()[-1])
print("value %.1fM e2e %.1fM ms %.1f"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"]), d["stage_ms"]); print(json.dumps(d["device_bvh_render"]))
PY
for w in "c3 32" "c2 0"; do
  set -- $w
  timeout 600 python bench.py --workload $1 --spp $2 --steps 2 --warmup 2 --no-cpu-baseline --no-bvh-microbench 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1 value %.1fM ms %.1f'%(d['value']/1e6,d['ms_per_step'])); print(json.dumps(d['device_bvh_render']))"
done
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r2v_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2v_pytest.log; tail -n 5 $O/r2v_pytest.log
