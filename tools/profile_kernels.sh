#!/bin/bash
# full ncu capture of selected kernels of the render bench: tools/profile_kernels.sh <tag> <regex> <skip> <count> [bench args]
set -u
TAG=$1; KREGEX=$2; SKIP=$3; COUNT=$4; shift 4
CMD="python bench.py --steps 1 --warmup 3 $*"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s $SKIP -c $COUNT -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
