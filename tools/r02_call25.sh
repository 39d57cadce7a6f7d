#!/bin/bash
# round 2, GPU call 25: tail hand-over limit re-swept on the faster wave kernels (C1, C2)
set -u
O=gpurun_out
export SLR_BENCH_AB=1
for T in 37888 9472 18944 75776 151552 37888; do
  SLRGPU_TAIL_PATHS=$T timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2I_c1_$T.json 2> $O/r2I_c1_$T.err
  SLRGPU_TAIL_PATHS=$T timeout 900 python bench.py --workload materials --spp 32 --steps 3 --warmup 3 > $O/r2I_c2_$T.json 2> $O/r2I_c2_$T.err
  for W in c1 c2; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2I_${W}_$T.json").read().strip().splitlines()[-1])
    print("$W tail<=$T", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), d["config"].get("tail_kernel"), d["config"].get("waves_per_frame"))
except Exception as e: print("$W $T", "ERR", e, open("$O/r2I_${W}_$T.err").read()[-400:])
PY
  done
done
