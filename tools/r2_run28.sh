#!/bin/bash
# round 2, GPU call 28 (2 GPUs): final bench under torchrun at N = 2, one process driving 2 devices, whole run over the device-built tree at N = 1, multi-GPU tests
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L | wc -l
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 ) > $O/bench_r2_final_n2.json 2> $O/bench_r2_final_n2.err; echo "N=2 ranks rc=$?"
( time timeout 600 python bench.py --gpus 2 --steps 2 --warmup 1 ) > $O/bench_r2_final_n2_threads.json 2> $O/bench_r2_final_n2_threads.err; echo "N=2 threads rc=$?"
( time timeout 600 python bench.py --tree device --steps 2 --warmup 1 --no-cpu-baseline --no-bvh-microbench ) > $O/bench_r2_final_device_tree.json 2> $O/bench_r2_final_device_tree.err; echo "device tree rc=$?"
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 ) > $O/bench_r2_final_ref_n2.json 2>/dev/null; echo "ref N=2 rc=$?"
python - <<PY
import json
for f in ("bench_r2_final_n2","bench_r2_final_n2_threads","bench_r2_final_device_tree","bench_r2_final_ref_n2"):
    try:
        d=json.loads(open("$O/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "n_gpus", d.get("n_gpus"), "value %.2fM"%(d["value"]/1e6), "e2e %.2fM"%(d["e2e"]["value"]/1e6), "ms %.1f"%d.get("ms_per_step",0), d.get("launch"), d.get("scaling"))
    except Exception as e:
        print(f, "no line", e)
PY
( timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) 2>&1 | tail -2
