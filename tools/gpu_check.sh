#!/bin/bash
# quick GPU round: parity tests, then stage timings on the Cornell+env, material-field and atrium scenes, then the BVH microbench
set -u
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== C2 cornell+env 1024^2 16spp"
python tools/render_once.py --scene 1 --res 1024 1024 --spp 16 --reps 3
echo "== C3 material field 1M tris 1920x1080 4spp"
python tools/render_once.py --scene 2 --res 1920 1080 --spp 4 --tris 1000000 --reps 3
echo "== C5 atrium 1920x1080 4spp"
python tools/render_once.py --scene 4 --res 1920 1080 --spp 4 --reps 3
echo "== C4 microbench"
python tools/microbench.py --all --iters 5
