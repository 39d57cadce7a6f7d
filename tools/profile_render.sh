#!/bin/bash
# ncu evidence for the render path (run under gpurun on one B200):
#   1. plain run (must exit 0)
#   2. every launch of one steady-state frame with its device time and DRAM bytes (-> launch list, traffic.json)
#   3. full captures of the first (full-pool) wave of a frame: one launch of each kernel family
# usage: tools/profile_render.sh <tag> <launches per frame> <launches per frame matching the capture regex> [extra bench args]
# (launches per frame = stats.kernel_launches of a frame; bench.py prints gpu_launches = steps x (that + 1 clear))
set -u
TAG=${1:-r1}; PER_FRAME=${2:-160}; MATCHED=${3:-112}; shift 3
CMD="python bench.py --steps 1 --warmup 3 $*"
# frames before the timed one: 1 plain + 1 profiled + 3 warm-up (the clear is a memset, not a kernel)
SKIP=$((5 * PER_FRAME))
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s $SKIP -c $PER_FRAME --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'materialKernel|extendKernel|surfaceKernel|shadowKernel|raygenKernel' -s $((3 * MATCHED)) -c 7 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
tail -1 gpurun_out/plain_$TAG.log | cut -c1-300
