#!/bin/bash
# tools/ncu_one.sh <tag> <mangled kernel regex> <skip> -- <command...>: plain run first, then one --set full capture
tag=$1; pat=$2; skip=$3; shift 4
"$@" > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$pat -s $skip -c 1 -f -o gpurun_out/prof_$tag "$@" > gpurun_out/ncu_$tag.log 2>&1
tail -n 1 gpurun_out/ncu_$tag.log
