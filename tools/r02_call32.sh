#!/bin/bash
# round 2, GPU call 32: results handed to the sink at the refill point in the flat renderer kernels too, now that a refill waits for 16 / 24 idle lanes
set -u
O=gpurun_out
export SLR_BENCH_AB=1
for L in libslrgpu.so libslrgpu_ds1.so libslrgpu.so libslrgpu_ds1.so; do
  SLRGPU_LIB=$L timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2N_c1_$L.json 2> $O/r2N_c1_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload materials --spp 32 --steps 3 --warmup 3 > $O/r2N_c2_$L.json 2> $O/r2N_c2_$L.err
  for W in c1 c2; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2N_${W}_$L.json").read().strip().splitlines()[-1])
    print("$W $L", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), d["config"].get("stage_ms_profiled_frame"))
except Exception as e: print("$W $L", "ERR", e, open("$O/r2N_${W}_$L.err").read()[-400:])
PY
  done
done
