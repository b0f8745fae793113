#!/bin/bash
# round 2, GPU call 3 (2 GPUs): multi-device tests, probe / parity tests, strong-scaling bench at N = 1, 2 (ranks) and 2 (threads)
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r2c_parity_report.jsonl
( PTRS_PARITY_REPORT=$PWD/$O/r2c_parity_report.jsonl timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -x -q -k "multi or comm or bxdf or lights or exact_shading or path_radiance" ) > $O/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2c_pytest.log; tail -n 15 $O/r2c_pytest.log
B="--steps 3 --warmup 1 --no-cpu-baseline --no-bvh-microbench"
timeout 900 python bench.py --gpus 1 $B > $O/r2c_bench_n1.json 2> $O/r2c_bench_n1.err; echo "n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $B > $O/r2c_bench_n2.json 2> $O/r2c_bench_n2.err; echo "n2 ranks rc=$?"
timeout 900 python bench.py --gpus 2 $B > $O/r2c_bench_n2_threads.json 2> $O/r2c_bench_n2_threads.err; echo "n2 threads rc=$?"
for f in n1 n2 n2_threads; do python - <<P
import json
try:
    d=json.load(open("$O/r2c_bench_$f.json"))
    print("$f", "value %.1fM e2e %.1fM ms %.1f e2e_ms %.1f"%(d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], d["e2e"]["ms_per_step"]), d["launch"], d["e2e"]["parts_last_step"])
except Exception as e:
    print("$f failed", e)
P
done
tail -n 5 $O/r2c_bench_n2.err $O/r2c_bench_n2_threads.err
