#!/bin/bash
# round 2, GPU call 15: A/B of the block-aggregated class append of the surface stage (128 / 256 threads), material at 4 blocks with the prefetch
set -u
O=gpurun_out
export SLR_BENCH_AB=1
for L in libslrgpu.so libslrgpu_nosb.so libslrgpu_sb256.so libslrgpu_mb4.so libslrgpu.so; do
  SLRGPU_LIB=$L timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2z_c1_$L.json 2> $O/r2z_c1_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload materials --spp 32 --steps 3 --warmup 3 > $O/r2z_c2_$L.json 2> $O/r2z_c2_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2z_c4_$L.json 2> $O/r2z_c4_$L.err
  for W in c1 c2 c4; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2z_${W}_$L.json").read().strip().splitlines()[-1])
    print("$W $L", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), d["config"].get("stage_ms_profiled_frame"))
except Exception as e: print("$W $L", "ERR", e, open("$O/r2z_${W}_$L.err").read()[-400:])
PY
  done
done
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
