#!/bin/bash
# round 2, GPU call 6: A/B of the deferred sink, the two-entry surface stage and the tail prefetch; parity suite on the result
set -u
mkdir -p gpurun_out
O=gpurun_out
for L in libslrgpu.so libslrgpu_nodefer.so libslrgpu_notailpf.so; do
  SLRGPU_LIB=$L timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2f_bench_c1_$L.json 2> $O/r2f_bench_c1_$L.err; echo "bench c1 $L rc=$?"
  SLRGPU_LIB=$L timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2f_bench_c4_$L.json 2> $O/r2f_bench_c4_$L.err; echo "bench c4 $L rc=$?"
  for W in c1 c4; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2f_bench_${W}_$L.json").read().strip().splitlines()[-1])
    print("$W $L", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), d["config"].get("stage_ms_profiled_frame"))
except Exception as e: print("$W $L", "ERR", e)
PY
  done
done
timeout 900 python bench.py --workload intersect --grid 2236 --rays 16777216 --steps 5 --warmup 3 --cpu-sample 200000 > $O/r2f_bench_c5.json 2> $O/r2f_bench_c5.err; python -c "
import json; d=json.loads(open('$O/r2f_bench_c5.json').read().strip().splitlines()[-1]); print('c5', round(d['value'],1), d['unit'])"
SLRGPU_LIB=libslrgpu_nodefer.so timeout 900 python bench.py --workload intersect --grid 2236 --rays 16777216 --steps 5 --warmup 3 --cpu-sample 200000 > $O/r2f_bench_c5_nodefer.json 2> $O/r2f_bench_c5_nodefer.err; python -c "
import json; d=json.loads(open('$O/r2f_bench_c5_nodefer.json').read().strip().splitlines()[-1]); print('c5 nodefer', round(d['value'],1), d['unit'])"
( timeout 2400 python -m pytest tests -m gpu -q -x > $O/r2f_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2f_gpu_tests.log ); tail -4 $O/r2f_gpu_tests.log
