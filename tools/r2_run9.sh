#!/bin/bash
# round 2, GPU call 9: pair parking / direct far / early parking A/B, parity of the traversal, exact-shading cost
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "intersect or counters or c4_hit or path_radiance or render_image or device_bvh or deeper or full_size" ) > $O/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2h_pytest.log; tail -n 6 $O/r2h_pytest.log
rm -f $O/r2h_ab.log
run() { # lib, label, env...
  lib=$1; label=$2; shift 2
  echo "=== $label" >> $O/r2h_ab.log
  env PTRS_B200_LIB=$PWD/$lib "$@" timeout 600 python tools/microbench.py --all --iters 4 2>&1 | grep -v "^tris=" | awk '{print $1,$2,$5,$6,$7,$8}' >> $O/r2h_ab.log
  for w in "c5 16" "c2 0" "c3 32"; do
    set -- $w "$@"; wl=$1; spp=$2; shift 2
    env PTRS_B200_LIB=$PWD/$lib "$@" timeout 600 python bench.py --workload $wl --spp $spp --steps 3 --warmup 2 --no-cpu-baseline --no-bvh-microbench 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$wl value %.1fM ms %.1f'%(d['value']/1e6,d['ms_per_step']), {k[3:]:round(v,1) for k,v in d['stage_ms'].items()})" >> $O/r2h_ab.log 2>&1
  done
}
run build/variants/libptrs_b200_base.so base
run build/variants/libptrs_b200_pair.so pair
run build/variants/libptrs_b200_pair_far.so pair_far
run build/variants/libptrs_b200_far_early.so far_early
run pathtracer_rs_b200/lib/libptrs_b200.so pair_far_early
run pathtracer_rs_b200/lib/libptrs_b200.so pair_far_early_boxmin16 PTRS_BOX_MIN=16
run pathtracer_rs_b200/lib/libptrs_b200.so pair_far_early_boxmin24 PTRS_BOX_MIN=24
cat $O/r2h_ab.log
python - <<P
import time, numpy as np
import pathtracer_rs_b200.gpu as gpu, pathtracer_rs_b200.host as host
for kind, res, spp, nt in ((host.SCENE_CORNELL_ENV, (1024, 1024), 64, 0), (host.SCENE_ATRIUM, (3840, 2160), 16, 262144), (host.SCENE_MATERIAL_FIELD, (1920, 1080), 32, 1000000)):
    flat, cam = host.make_scene(kind, seed=1, n_tris=nt, res=res, env_hdr=host.TANK_FARM_HDR if kind == host.SCENE_CORNELL_ENV else None)
    scene = gpu.RenderScene(flat)
    integ = gpu.PathIntegrator(gpu.SamplerBuilder(spp), max_depth=15)
    film = gpu.Film(cam.width, cam.height)
    for exact in (False, True, False, True):
        film.clear()
        st = integ.render(cam, scene, film, exact_shading=exact)
        print("scene", kind, "exact" if exact else "default", "shade ms %.1f total %.1f" % (st["ms_shade"], st["ms_total"]))
    scene.close()
P
