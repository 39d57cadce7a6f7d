#!/bin/bash
# round 2, GPU call 34: steps between refill checks by kernel (instanced 2, batch 2, flat renderer 1) -- bit-exact tests, the
# bench lines, and 3 steps for the instanced kernels
set -u
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "intersect or occlu or traversal or sbvh or nested or instanc or motion" 2>&1 | tail -3
export SLR_BENCH_AB=1
for L in libslrgpu.so libslrgpu_si3.so libslrgpu.so; do
  SLRGPU_LIB=$L timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2Q_c4_$L.json 2> $O/r2Q_c4_$L.err
  SLRGPU_LIB=$L timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2Q_c1_$L.json 2> $O/r2Q_c1_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload intersect --steps 5 --warmup 3 --cpu-sample 20000 > $O/r2Q_c5_$L.json 2> $O/r2Q_c5_$L.err
  for W in c4 c1 c5; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2Q_${W}_$L.json").read().strip().splitlines()[-1])
    print("$W $L", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), d["config"].get("stage_ms_profiled_frame"))
except Exception as e: print("$W $L", "ERR", e, open("$O/r2Q_${W}_$L.err").read()[-400:])
PY
  done
done
