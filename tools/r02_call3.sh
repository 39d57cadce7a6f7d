#!/bin/bash
# round 2, GPU call 3: full parity suite (drop-in binary, textured scene, multi-replica), instanced-walk register A/B
set -u
mkdir -p gpurun_out
O=gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q > $O/r2c_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2c_gpu_tests.log )
tail -6 $O/r2c_gpu_tests.log
for L in libslrgpu.so libslrgpu_norcp.so libslrgpu_rcp84.so; do
  SLRGPU_LIB=$L timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2c_bench_c4_$L.json 2> $O/r2c_bench_c4_$L.err; echo "bench c4 $L rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2c_bench_c4_$L.json").read().strip().splitlines()[-1])
    print("$L", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), d["config"].get("stage_ms_profiled_frame"))
except Exception as e: print("$L", "ERR", e)
PY
done
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2c_bench_c1.json 2> $O/r2c_bench_c1.err; echo "bench c1 rc=$?"; cut -c1-200 $O/r2c_bench_c1.json
