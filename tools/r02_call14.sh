#!/bin/bash
# round 2, GPU call 14: A/B of the queue prefetch hints (stage kernels, walk refill, shadow sink) and the two-wide slab arithmetic
set -u
O=gpurun_out
export SLR_BENCH_AB=1
for L in libslrgpu.so libslrgpu_pfs.so libslrgpu_pfw.so libslrgpu_pf.so libslrgpu_pf2.so libslrgpu_ps.so libslrgpu_all.so libslrgpu.so; do
  SLRGPU_LIB=$L timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2y_c1_$L.json 2> $O/r2y_c1_$L.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2y_c1_$L.json").read().strip().splitlines()[-1])
    print("c1 $L", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), d["config"].get("stage_ms_profiled_frame"))
except Exception as e: print("c1 $L", "ERR", e, open("$O/r2y_c1_$L.err").read()[-400:])
PY
done
for L in libslrgpu.so libslrgpu_pf2.so libslrgpu_ps.so libslrgpu_all.so; do
  SLRGPU_LIB=$L timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2y_c4_$L.json 2> $O/r2y_c4_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload materials --spp 32 --steps 3 --warmup 3 > $O/r2y_c2_$L.json 2> $O/r2y_c2_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload intersect --steps 5 --warmup 3 --cpu-sample 20000 > $O/r2y_c5_$L.json 2> $O/r2y_c5_$L.err
  for W in c4 c2 c5; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2y_${W}_$L.json").read().strip().splitlines()[-1])
    print("$W $L", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), d["config"].get("stage_ms_profiled_frame"))
except Exception as e: print("$W $L", "ERR", e, open("$O/r2y_${W}_$L.err").read()[-400:])
PY
  done
done
SLRGPU_LIB=libslrgpu_all.so timeout 1500 python -m pytest tests -m gpu -q -x -k "intersect or sbvh or render or bpt" 2>&1 | tail -3
