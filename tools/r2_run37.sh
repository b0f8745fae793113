#!/bin/bash
# round 2, GPU call 37: bench lines at HEAD (default with CPU baseline + microbench, reference arm, C1 / C2 / C3, device-built tree), then ncu --set full of the traversal kernels
set -u
O=gpurun_out
mkdir -p $O
( time timeout 900 python bench.py ) > $O/bench_r2_final3.json 2> $O/bench_r2_final3.err; echo "bench rc=$?"; tail -3 $O/bench_r2_final3.err
( timeout 600 python bench.py --impl reference --steps 1 --warmup 0 ) > $O/bench_r2_final3_reference.json 2>/dev/null; echo "ref rc=$?"
for w in c1 c2 c3; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-bvh-microbench > $O/bench_r2_final3_$w.json 2>/dev/null; echo "$w rc=$?"
done
timeout 600 python bench.py --tree device --steps 2 --warmup 2 --no-cpu-baseline --no-bvh-microbench > $O/bench_r2_final3_device_tree.json 2>/dev/null; echo "device tree rc=$?"
python - <<PY
import json
for f in ("bench_r2_final3","bench_r2_final3_reference","bench_r2_final3_c1","bench_r2_final3_c2","bench_r2_final3_c3","bench_r2_final3_device_tree"):
    try:
        d=json.loads(open("$O/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.2fM"%(d["value"]/1e6), "e2e %.2fM"%(d["e2e"]["value"]/1e6), "ms %.1f"%d.get("ms_per_step",0), d.get("clocks"), {k:round(v,1) for k,v in d.get("stage_ms",{}).items()})
    except Exception as e:
        print(f, "no line", e)
PY
TAG=r2g
CMD="python tools/render_once.py --scene 4 --res 3840 2160 --spp 2 --tris 262144 --reps 1"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:extend_kernelILb0ELb0 -s 1 -c 1 -f -o $O/prof_c5_extend_$TAG $CMD > $O/ncu_c5_extend_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:connect_kernelILb0ELb0 -s 1 -c 1 -f -o $O/prof_c5_connect_$TAG $CMD > $O/ncu_c5_connect_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:intersect_kernelILb0ELb0ELb0 -s 1 -c 1 -f -o $O/prof_intersect_$TAG python tools/microbench.py --rays incoherent --iters 2 > $O/ncu_intersect_$TAG.log 2>&1
for f in $O/ncu_c5_*_$TAG.log $O/ncu_intersect_$TAG.log; do tail -n 1 $f; done
