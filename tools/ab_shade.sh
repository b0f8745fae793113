#!/bin/bash
# parity tests, then bench lines of C2 / C3 / C5 for an A/B of a shade change:  tools/ab_shade.sh <tag> [skip-tests]
set -u
tag=${1:-x}
if [ "${2:-}" != "skip-tests" ]; then timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; fi
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bvh-microbench > gpurun_out/bench_c2_$tag.json 2> gpurun_out/bench_c2_$tag.err
python bench.py --workload c3 --spp 32 --steps 3 --warmup 3 --no-bvh-microbench --no-cpu-baseline > gpurun_out/bench_c3_$tag.json 2> gpurun_out/bench_c3_$tag.err
python bench.py --workload c5 --spp 16 --steps 3 --warmup 3 --no-bvh-microbench --no-cpu-baseline > gpurun_out/bench_c5_$tag.json 2> gpurun_out/bench_c5_$tag.err
python - "$tag" <<'PY'
import json,glob,sys
for f in sorted(glob.glob('gpurun_out/bench_c*_%s.json' % sys.argv[1])):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f,'ERR', open(f.replace('.json','.err')).read()[-800:]); continue
    print(f, 'value %.1fM e2e %.1fM ms %.1f'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step']), {k:round(v,1) for k,v in d['stage_ms'].items()})
PY
