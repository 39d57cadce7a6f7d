#!/bin/bash
# round 2, GPU call 35: the committed library (instanced walk: three steps between refill checks) -- whole parity suite, C4 line
set -u
O=gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q > $O/r2z_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2z_gpu_tests.log ); tail -3 $O/r2z_gpu_tests.log
timeout 600 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2z_bench_c4_spp16.json 2> $O/r2z_bench_c4_spp16.err; echo "c4 rc=$?"
python - <<PY
import json
d=json.loads(open("$O/r2z_bench_c4_spp16.json").read().strip().splitlines()[-1])
print("c4", round(d["value"],2), d["unit"], "e2e", round(d["e2e"]["value"],2), "cpu", (d.get("cpu_baseline") or {}).get("value"), "parity", (d.get("image_parity") or {}).get("ratio_to_floor"))
PY
