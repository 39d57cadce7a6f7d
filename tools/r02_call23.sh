#!/bin/bash
# round 2, GPU call 23: entries that see emission queued by the surface stage and added by emissionKernel with full warps
set -u
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "render or multi or debug or save or dropin or scenes or probe" 2>&1 | tail -4
export SLR_BENCH_AB=1
for L in libslrgpu.so libslrgpu.so; do
  SLRGPU_LIB=$L timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2G_c1_$L.json 2> $O/r2G_c1_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload materials --spp 32 --steps 3 --warmup 3 > $O/r2G_c2_$L.json 2> $O/r2G_c2_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload ibl --spp 32 --steps 3 --warmup 3 > $O/r2G_c3_$L.json 2> $O/r2G_c3_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2G_c4_$L.json 2> $O/r2G_c4_$L.err
  for W in c1 c2 c3 c4; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2G_${W}_$L.json").read().strip().splitlines()[-1])
    print("$W $L", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), d["config"].get("stage_ms_profiled_frame"))
except Exception as e: print("$W $L", "ERR", e, open("$O/r2G_${W}_$L.err").read()[-400:])
PY
  done
done
