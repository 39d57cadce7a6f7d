#!/bin/bash
# round 2, GPU call 30: tail hand-over limit by scene (C1 / C3 / C4 at 75 776 paths), any-hit walk refilling at 24 idle lanes
set -u
O=gpurun_out
export SLR_BENCH_AB=1
run() { # tag lib tail workload args
  local tag=$1 lib=$2 tail=$3; shift 3
  SLRGPU_LIB=$lib SLRGPU_TAIL_PATHS=$tail timeout 900 python bench.py "$@" > $O/r2M_$tag.json 2> $O/r2M_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2M_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), d["config"].get("tail_kernel"), d["config"].get("stage_ms_profiled_frame"))
except Exception as e: print("$tag", "ERR", e, open("$O/r2M_$tag.err").read()[-300:])
PY
}
for T in 37888 75776; do
  run c1_t$T libslrgpu.so $T --steps 10 --warmup 3
  run c3_t$T libslrgpu.so $T --workload ibl --spp 64 --steps 3 --warmup 3
  run c4_t$T libslrgpu.so $T --workload instanced --spp 16 --steps 3 --warmup 3
done
run c1_sh24 libslrgpu_sh24.so 37888 --steps 10 --warmup 3
run c2_sh24 libslrgpu_sh24.so 37888 --workload materials --spp 32 --steps 3 --warmup 3
run c2_base libslrgpu.so 37888 --workload materials --spp 32 --steps 3 --warmup 3
run c1_base libslrgpu.so 37888 --steps 10 --warmup 3
