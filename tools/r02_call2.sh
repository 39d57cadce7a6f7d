#!/bin/bash
# round 2, GPU call 2: full parity suite on the new pipeline, A/B of the loop driver and the prefetch variant
set -u
mkdir -p gpurun_out
O=gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q > $O/r2b_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2b_gpu_tests.log )
tail -4 $O/r2b_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2b_bench_c1.json 2> $O/r2b_bench_c1.err; echo "bench c1 rc=$?"
SLRGPU_HOST_LOOP=1 timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2b_bench_c1_hostloop.json 2> $O/r2b_bench_c1_hostloop.err; echo "bench c1 hostloop rc=$?"
timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2b_bench_c4_spp16.json 2> $O/r2b_bench_c4.err; echo "bench c4 rc=$?"
timeout 900 python bench.py --workload materials --spp 32 --steps 3 --warmup 3 > $O/r2b_bench_c2_spp32.json 2> $O/r2b_bench_c2.err; echo "bench c2 rc=$?"
timeout 900 python bench.py --workload intersect --grid 2236 --rays 16777216 --steps 5 --warmup 3 --cpu-sample 200000 > $O/r2b_bench_c5.json 2> $O/r2b_bench_c5.err; echo "bench c5 rc=$?"
SLRGPU_LIB=libslrgpu_pf.so timeout 900 python bench.py --workload intersect --grid 2236 --rays 16777216 --steps 5 --warmup 3 --cpu-sample 200000 > $O/r2b_bench_c5_pf.json 2> $O/r2b_bench_c5_pf.err; echo "bench c5 pf rc=$?"
SLRGPU_LIB=libslrgpu_pf.so timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2b_bench_c4_spp16_pf.json 2> $O/r2b_bench_c4_pf.err; echo "bench c4 pf rc=$?"
for f in c1 c1_hostloop c4_spp16 c4_spp16_pf c2_spp32 c5 c5_pf; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2b_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), d["config"].get("stage_ms_profiled_frame"))
except Exception as e: print("$f", "ERR", e)
PY
done
