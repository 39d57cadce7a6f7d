#!/bin/bash
# ncu evidence for the render path (run under gpurun on one B200):
#   1. plain run (must exit 0)   2. every launch with its device time   3. full capture of the top kernel
# usage: tools/profile_render.sh <tag> <kernel-regex> [extra bench args]
set -u
TAG=${1:-r1}; KREGEX=${2:-shadeKernel}; shift 2
CMD="python bench.py --size 256 --spp 16 --steps 2 --warmup 3 $*"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 8 -c 3 -f -o gpurun_out/prof_${TAG}_$KREGEX $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
tail -2 gpurun_out/plain_$TAG.log | cut -c1-600
