#!/bin/bash
# Profiling pass of a round, run under gpurun.  usage: tools/ncu_round.sh <tag>
#  0. the plain bench line (no profiler)
#  1. launch list (gpu__time_duration) of the bench command itself, CPU baseline and microbench legs switched off
#  2. DRAM bytes of every extend launch of a 64-spp render (metrics-only pass) -> roofline.traffic
#  3. full captures (--set full, source) of extend / shade<Matte> / connect / connect_resolve of a 16-spp render, of the
#     standalone closest-hit kernel on incoherent rays over the 10 M-triangle terrain, and of the BVH refit kernel
set -u
TAG=${1:-r1s3}
O=gpurun_out
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-bvh-microbench"
$BENCH > $O/plain_bench_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_bench_$TAG.csv $BENCH > $O/ncu_list_$TAG.log 2>&1
CMD64="python tools/render_once.py --scene 1 --res 1024 1024 --spp 64 --reps 1"
$CMD64 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base mangled -k regex:extend_kernelILb0 -c 400 --csv --log-file $O/extend_dram_$TAG.csv $CMD64 > $O/ncu_extdram_$TAG.log 2>&1
CMD="python tools/render_once.py --scene 1 --res 1024 1024 --spp 16 --reps 1"
for k in extend_kernelILb0:extend shade_kernelILi0:shade0 connect_kernelILb0:connect connect_resolve:resolve; do
  pat=${k%%:*}; name=${k##*:}
  $CMD > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$pat -s 2 -c 1 -f -o $O/prof_${name}_$TAG $CMD > $O/ncu_${name}_$TAG.log 2>&1
done
CMD2="python tools/microbench.py --rays incoherent --iters 2"
$CMD2 > $O/plain_micro_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:intersect_kernelILb0ELb0 -s 1 -c 1 -f -o $O/prof_intersect_$TAG $CMD2 > $O/ncu_intersect_$TAG.log 2>&1
CMD3="python tools/bvh_build_time.py --reps 2"
$CMD3 > $O/plain_bvh_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:refit_kernel -s 1 -c 1 -f -o $O/prof_refit_$TAG $CMD3 > $O/ncu_refit_$TAG.log 2>&1
$CMD3 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $O/bvh_launches_$TAG.csv python tools/bvh_build_time.py --reps 1 > $O/ncu_bvhlist_$TAG.log 2>&1
for f in $O/ncu_*_$TAG.log; do tail -n 1 $f; done
cat $O/plain_bvh_$TAG.log
