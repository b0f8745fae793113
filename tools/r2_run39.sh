#!/bin/bash
# round 2, GPU call 39 (2 GPUs): bench at HEAD under torchrun at N = 2 (strong scaling check after the traversal changes)
set -u
O=gpurun_out
mkdir -p $O
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 3 --warmup 2 ) > $O/bench_r2_final3_n2.json 2> $O/bench_r2_final3_n2.err; echo "N=2 ranks rc=$?"
python - <<PY
import json
d=json.loads(open("$O/bench_r2_final3_n2.json").read().strip().splitlines()[-1])
print("n_gpus", d.get("n_gpus"), "value %.2fM"%(d["value"]/1e6), "e2e %.2fM"%(d["e2e"]["value"]/1e6), "ms %.1f"%d.get("ms_per_step",0), d.get("launch"), d.get("scaling"), d.get("clocks"))
PY
