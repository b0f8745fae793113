#!/bin/bash
# stage times of C2 / C3 / C5 for the default library and every build/variants/libptrs_b200_*.so
set -u
for lib in default build/variants/libptrs_b200_*.so; do
  if [ $lib = default ]; then unset PTRS_B200_LIB; else export PTRS_B200_LIB=$PWD/$lib; fi
  for w in "c2 0" "c3 32" "c5 16"; do
    set -- $w
    python bench.py --workload $1 --spp $2 --steps 3 --warmup 3 --no-cpu-baseline --no-bvh-microbench 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib $1', 'value %.1fM ms %.1f'%(d['value']/1e6,d['ms_per_step']), {k[3:]:round(v,1) for k,v in d['stage_ms'].items()})"
  done
done
