#!/bin/bash
# ncu evidence for the render path (run under gpurun on one B200):
#   1. plain run (must exit 0)   2. every launch of one steady-state frame with its device time
#   3. full captures of the top kernels (material<Lambert>, extend, surface)
# usage: tools/profile_render.sh <tag> [extra bench args]
set -u
TAG=${1:-r1}; shift 1
CMD="python bench.py --steps 1 --warmup 3 $*"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 4500 -c 1150 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'materialKernel|extendKernel|surfaceKernel|shadowKernel|raygenKernel' -s 60 -c 7 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
tail -1 gpurun_out/plain_$TAG.log | cut -c1-400
