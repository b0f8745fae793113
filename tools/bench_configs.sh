#!/bin/bash
# Bench lines of the other BASELINE configs (C1, C3, C5) + launch lists of C3 and C5 for the per-kernel shares.
# C5 at N=1 is run at reduced spp (the full 1024 spp is the 8-GPU job); the line says so.
set -u
tag=${1:-r1s4}
mkdir -p gpurun_out
python bench.py --workload c1 --steps 3 --warmup 3 --no-bvh-microbench > gpurun_out/bench_c1_$tag.json 2> gpurun_out/bench_c1_$tag.err
python bench.py --workload c3 --spp 32 --steps 3 --warmup 3 --no-bvh-microbench > gpurun_out/bench_c3_$tag.json 2> gpurun_out/bench_c3_$tag.err
python bench.py --workload c5 --spp 16 --steps 3 --warmup 3 --no-bvh-microbench > gpurun_out/bench_c5_$tag.json 2> gpurun_out/bench_c5_$tag.err
for c in c3 c5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${c}_$tag.csv \
    python bench.py --workload $c --spp 4 --steps 1 --warmup 0 --no-bvh-microbench --no-cpu-baseline > gpurun_out/ncu_${c}_$tag.log 2>&1
done
tail -c 600 gpurun_out/bench_c3_$tag.json
