#!/bin/bash
# ncu evidence for the render path (run under gpurun on one B200):
#   1. plain run (must exit 0)
#   2. every launch of one steady-state frame with its device time and DRAM bytes (-> launch list, traffic.json)
#   3. full captures of one launch of each kernel
# usage: tools/profile_render.sh <tag> <launches per frame> [extra bench args]
set -u
TAG=${1:-r1}; PER_FRAME=${2:-955}; shift 2
CMD="python bench.py --steps 1 --warmup 3 $*"
SKIP=$((4 * PER_FRAME + 20))
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s $SKIP -c $PER_FRAME --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'materialKernel|extendKernel|surfaceKernel|shadowKernel|raygenKernel' -s $((PER_FRAME + 60)) -c 7 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
tail -1 gpurun_out/plain_$TAG.log | cut -c1-300
