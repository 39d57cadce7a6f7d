#!/bin/bash
# round 2, GPU call 27: refills that book their queue entries and hint them into L1 two / four steps ahead
set -u
O=gpurun_out
SLRGPU_LIB=libslrgpu_ra2.so timeout 1200 python -m pytest tests -m gpu -q -x -k "intersect or occlu or traversal or sbvh" 2>&1 | tail -3
export SLR_BENCH_AB=1
for L in libslrgpu.so libslrgpu_ra2.so libslrgpu_ra4.so libslrgpu.so; do
  SLRGPU_LIB=$L timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2L_c1_$L.json 2> $O/r2L_c1_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload materials --spp 32 --steps 3 --warmup 3 > $O/r2L_c2_$L.json 2> $O/r2L_c2_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2L_c4_$L.json 2> $O/r2L_c4_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload intersect --steps 5 --warmup 3 --cpu-sample 20000 > $O/r2L_c5_$L.json 2> $O/r2L_c5_$L.err
  for W in c1 c2 c4 c5; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2L_${W}_$L.json").read().strip().splitlines()[-1])
    print("$W $L", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), d["config"].get("stage_ms_profiled_frame"))
except Exception as e: print("$W $L", "ERR", e, open("$O/r2L_${W}_$L.err").read()[-400:])
PY
  done
done
