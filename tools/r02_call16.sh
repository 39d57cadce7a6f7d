#!/bin/bash
# round 2, GPU call 16: the branch-free node step (traverse.cuh walkNode) -- bit-exact hit tests first, then A/B against the
# previous walk (libslrgpu_oldwalk.so) and of the three class-append forms of the surface stage
set -u
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "intersect or sbvh or occlu or traversal or nested or probe" 2>&1 | tail -5
timeout 600 python -m pytest "tests/test_dropin.py" -m gpu -q -x -k "nested or motion" 2>&1 | tail -30
export SLR_BENCH_AB=1
for L in libslrgpu_oldwalk.so libslrgpu_sa0.so libslrgpu_sa1.so libslrgpu.so; do
  SLRGPU_LIB=$L timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2A_c1_$L.json 2> $O/r2A_c1_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload materials --spp 32 --steps 3 --warmup 3 > $O/r2A_c2_$L.json 2> $O/r2A_c2_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2A_c4_$L.json 2> $O/r2A_c4_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload intersect --steps 5 --warmup 3 --cpu-sample 20000 > $O/r2A_c5_$L.json 2> $O/r2A_c5_$L.err
  for W in c1 c2 c4 c5; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2A_${W}_$L.json").read().strip().splitlines()[-1])
    print("$W $L", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), d["config"].get("stage_ms_profiled_frame"))
except Exception as e: print("$W $L", "ERR", e, open("$O/r2A_${W}_$L.err").read()[-400:])
PY
  done
done
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
