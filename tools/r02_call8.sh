#!/bin/bash
# round 2, GPU call 8 (1 GPU): the record of the final pipeline -- parity suite, bench lines of all five BASELINE configs,
# C1 launch list (one frame between cudaProfilerStart/Stop) and first-wave full capture
set -u
mkdir -p gpurun_out
O=gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q > $O/r2h_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2h_gpu_tests.log ); tail -4 $O/r2h_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2h_bench_c1.json 2> $O/r2h_bench_c1.err; echo "c1 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2h_bench_c1_ref.json 2> $O/r2h_bench_c1_ref.err; echo "c1 ref rc=$?"
timeout 900 python bench.py --workload materials --steps 3 --warmup 3 > $O/r2h_bench_c2.json 2> $O/r2h_bench_c2.err; echo "c2 rc=$?"
timeout 900 python bench.py --workload ibl --steps 3 --warmup 3 > $O/r2h_bench_c3.json 2> $O/r2h_bench_c3.err; echo "c3 rc=$?"
timeout 1200 python bench.py --workload intersect --grid 2236 --rays 67108864 --steps 5 --warmup 3 --cpu-sample 2000000 > $O/r2h_bench_c5.json 2> $O/r2h_bench_c5.err; echo "c5 rc=$?"
for f in c1 c2 c3 c5; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2h_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None, "roofline", d["roofline"]["bound"], round(d["roofline"]["frac"],3), d.get("image_parity"))
except Exception as e: print("$f", "ERR", e)
PY
done
NCU_L="ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
timeout 900 $NCU_L --log-file $O/r2h_launches_c1.csv python tools/ncu_frame.py --workload cornell_spheres > $O/r2h_ncu_c1.log 2>&1; echo "launch list c1 rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'materialKernel|extendKernel|surfaceKernel|shadowKernel|raygenKernel' -c 7 -f -o $O/r2h_prof_c1 python tools/ncu_frame.py --workload cornell_spheres > $O/r2h_ncu_full_c1.log 2>&1; echo "full c1 rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tailKernel -c 8 -f -o $O/r2h_prof_tail python tools/ncu_frame.py --workload cornell_spheres > $O/r2h_ncu_full_tail.log 2>&1; echo "full tail rc=$?"
