#!/bin/bash
# round 2, GPU call 33: two walk steps between refill checks (long instanced walks: C4), against one
set -u
O=gpurun_out
export SLR_BENCH_AB=1
for L in libslrgpu.so libslrgpu_s2.so libslrgpu.so libslrgpu_s2.so; do
  SLRGPU_LIB=$L timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2P_c4_$L.json 2> $O/r2P_c4_$L.err
  SLRGPU_LIB=$L timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2P_c1_$L.json 2> $O/r2P_c1_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload intersect --steps 5 --warmup 3 --cpu-sample 20000 > $O/r2P_c5_$L.json 2> $O/r2P_c5_$L.err
  for W in c4 c1 c5; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2P_${W}_$L.json").read().strip().splitlines()[-1])
    print("$W $L", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), d["config"].get("stage_ms_profiled_frame"))
except Exception as e: print("$W $L", "ERR", e, open("$O/r2P_${W}_$L.err").read()[-400:])
PY
  done
done
