#!/bin/bash
# round 2, GPU call 31: the committed library once more -- whole parity suite, smoke(), the driver's bench line, C2 / C4 lines
set -u
O=gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q > $O/r2z_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2z_gpu_tests.log ); tail -3 $O/r2z_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > $O/r2z_bench_c1.json 2> $O/r2z_bench_c1.err; echo "c1 rc=$?"
timeout 900 python bench.py --workload materials --steps 3 --warmup 3 --parity-paths 3e8 > $O/r2z_bench_c2.json 2> $O/r2z_bench_c2.err; echo "c2 rc=$?"
timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2z_bench_c4_spp16.json 2> $O/r2z_bench_c4_spp16.err; echo "c4 rc=$?"
for f in c1 c2 c4_spp16; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2z_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],2), "cpu", (d.get("cpu_baseline") or {}).get("value"), "parity", (d.get("image_parity") or {}).get("ratio_to_floor"), d["config"].get("tail_kernel"))
except Exception as e: print("$f", "ERR", e, open("$O/r2z_bench_$f.err").read()[-300:])
PY
done
