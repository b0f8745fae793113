#!/bin/bash
# round 2, GPU call 20: launch list of the PLOC build (10 M triangles)
set -u
O=gpurun_out
mkdir -p $O
PTRS_BVH_BUILDER=ploc timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv --log-file $O/r2t_ploc_launches.csv python tools/bvh_build_time.py --reps 1 > $O/r2t_ncu.log 2>&1
tail -3 $O/r2t_ncu.log
python tools/launch_summary.py $O/r2t_ploc_launches.csv | cut -c1-160 | head -30
