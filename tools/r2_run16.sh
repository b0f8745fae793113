#!/bin/bash
# round 2, GPU call 16 (8 GPUs): strong scaling of the default bench (4K atrium, 128 spp per step) at N = 8, 4, 2, 1 the way
# the driver launches it, then one process driving 8 devices through ptrs_multi_render, then the multi-GPU tests
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L | wc -l
for N in 8 4 2 1; do
  if [ "$N" -gt 1 ]; then
    L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
  else
    L="python"
  fi
  ( time timeout 900 $L bench.py --gpus $N --steps 3 --warmup 3 --no-bvh-microbench --no-cpu-baseline ) > $O/bench_scale_n$N.json 2> $O/bench_scale_n$N.err
  echo "N=$N rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/bench_scale_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value %.1fM e2e %.1fM ms %.1f"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"]), d["e2e"].get("parts_last_step"))
except Exception as e:
    print("N=$N no line", e)
PY
done
( time timeout 900 python bench.py --gpus 8 --steps 3 --warmup 3 --no-bvh-microbench --no-cpu-baseline ) > $O/bench_scale_n8_threads.json 2> $O/bench_scale_n8_threads.err
echo "threads N=8 rc=$?"; tail -c 600 $O/bench_scale_n8_threads.json
( timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > $O/r2n_pytest_multi.log 2>&1; tail -n 3 $O/r2n_pytest_multi.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 ) > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref N=2 rc=$?"; tail -c 400 $O/bench_ref_n2.json
