#!/bin/bash
# bench.py at N = 1, 2, 4, 8 on one box (the driver's launch line), plus the reference arm once.
# usage: tools/scaling_run.sh <tag> [list of N, default "1 2 4 8"]
TAG=${1:-r1}; NS=${2:-"1 2 4 8"}
mkdir -p gpurun_out
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/scale_${TAG}_ref.json 2> gpurun_out/scale_err.log
FILES=gpurun_out/scale_${TAG}_ref.json
for N in $NS; do
  if [ $N -eq 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/scale_${TAG}_n$N.json 2>> gpurun_out/scale_err.log
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_${TAG}_n$N.json 2>> gpurun_out/scale_err.log
  fi
  echo "N=$N rc=$?"
  FILES="$FILES gpurun_out/scale_${TAG}_n$N.json"
done
python tools/bench_summary.py $FILES | grep -v "stages\|clocks\|roofline\|traversal"
