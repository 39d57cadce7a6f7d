#!/bin/bash
# round 2, GPU call 24: leaf records of a step tested by the whole warp (flat walk), taken when the longest list holds >= 3 / >= 6 records
set -u
O=gpurun_out
for L in libslrgpu_coop3.so; do
  SLRGPU_LIB=$L timeout 1200 python -m pytest tests -m gpu -q -x -k "intersect or occlu or traversal or sbvh" 2>&1 | tail -4
done
export SLR_BENCH_AB=1
for L in libslrgpu.so libslrgpu_coop3.so libslrgpu_coop6.so libslrgpu.so; do
  SLRGPU_LIB=$L timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2H_c1_$L.json 2> $O/r2H_c1_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload materials --spp 32 --steps 3 --warmup 3 > $O/r2H_c2_$L.json 2> $O/r2H_c2_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload intersect --steps 5 --warmup 3 --cpu-sample 20000 > $O/r2H_c5_$L.json 2> $O/r2H_c5_$L.err
  for W in c1 c2 c5; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2H_${W}_$L.json").read().strip().splitlines()[-1])
    print("$W $L", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), d["config"].get("stage_ms_profiled_frame"))
except Exception as e: print("$W $L", "ERR", e, open("$O/r2H_${W}_$L.err").read()[-400:])
PY
  done
done
