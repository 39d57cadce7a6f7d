#!/bin/bash
# round 2, GPU call 26: refill threshold of the walk (idle lanes that trigger a refill) re-swept on the leaner node step
set -u
O=gpurun_out
export SLR_BENCH_AB=1
for L in libslrgpu_ri16.so libslrgpu_ri20.so libslrgpu_ri24.so libslrgpu_ri12.so; do
  SLRGPU_LIB=$L timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2K_c1_$L.json 2> $O/r2K_c1_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2K_c4_$L.json 2> $O/r2K_c4_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload materials --spp 32 --steps 3 --warmup 3 --cpu-sample 20000 > $O/r2K_c2_$L.json 2> $O/r2K_c5_$L.err
  for W in c1 c4 c2; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2K_${W}_$L.json").read().strip().splitlines()[-1])
    print("$W $L", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), d["config"].get("stage_ms_profiled_frame"))
except Exception as e: print("$W $L", "ERR", e, open("$O/r2K_${W}_$L.err").read()[-400:])
PY
  done
done
