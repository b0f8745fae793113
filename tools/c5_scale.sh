#!/bin/bash
# BASELINE configs[4] at full size: 3840x2160, 1024 spp sharded over 8 GPUs (128 spp each), film reduced with NCCL
set -u
N=${1:-8}; SPP=${2:-128}; tag=${3:-r1s4}
if [ "$N" -gt 1 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload c5 --spp $SPP --steps 2 --warmup 3 --no-bvh-microbench --no-cpu-baseline > gpurun_out/bench_c5_${N}gpu_$tag.json 2> gpurun_out/bench_c5_${N}gpu_$tag.err
else
  python bench.py --gpus 1 --workload c5 --spp $SPP --steps 2 --warmup 3 --no-bvh-microbench --no-cpu-baseline > gpurun_out/bench_c5_${N}gpu_$tag.json 2> gpurun_out/bench_c5_${N}gpu_$tag.err
fi
tail -c 1500 gpurun_out/bench_c5_${N}gpu_$tag.json; tail -3 gpurun_out/bench_c5_${N}gpu_$tag.err
