#!/bin/bash
# round 2, GPU call 13 (same programme as call 5, at the final traversal kernel): the whole GPU suite at HEAD with the parity report, the default bench line, then the profiling pass of the
# default workload (C5 4K atrium): launch list of the bench command, DRAM bytes of every traversal launch, full captures of the
# second-bounce extend / connect launches and of shade<Matte>; and of the incoherent closest-hit kernel on the 10 M-triangle tree
set -u
TAG=r2f
O=gpurun_out
mkdir -p $O
rm -f $O/${TAG}_parity_report.jsonl
( time PTRS_PARITY_REPORT=$PWD/$O/${TAG}_parity_report.jsonl timeout 1500 python -m pytest tests -m gpu -q ) > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest.log; tail -n 6 $O/${TAG}_pytest.log
( time timeout 1200 python bench.py --steps 5 --warmup 3 ) > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
BENCH="python bench.py --spp 16 --steps 1 --warmup 0 --no-cpu-baseline --no-bvh-microbench"
$BENCH > $O/plain_bench_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_bench_$TAG.csv $BENCH > $O/ncu_list_$TAG.log 2>&1
CMD="python tools/render_once.py --scene 4 --res 3840 2160 --spp 4 --tris 262144 --reps 1"
$CMD > $O/plain_c5_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base mangled -k regex:'(extend|connect)_kernelILb0' -c 200 --csv --log-file $O/traversal_dram_$TAG.csv $CMD > $O/ncu_travdram_$TAG.log 2>&1
for k in extend_kernelILb0:extend connect_kernelILb0:connect shade_kernelILi0:shade0; do
  pat=${k%%:*}; name=${k##*:}
  $CMD > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$pat -s 1 -c 1 -f -o $O/prof_c5_${name}_$TAG $CMD > $O/ncu_c5_${name}_$TAG.log 2>&1
done
CMD2="python tools/microbench.py --rays incoherent --iters 2"
$CMD2 > $O/plain_micro_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:intersect_kernelILb0ELb0 -s 1 -c 1 -f -o $O/prof_intersect_$TAG $CMD2 > $O/ncu_intersect_$TAG.log 2>&1
for f in $O/ncu_*_$TAG.log; do echo $f; tail -n 1 $f; done
ls -la $O/*.ncu-rep 2>/dev/null
