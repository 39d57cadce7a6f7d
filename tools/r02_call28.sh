#!/bin/bash
# round 2, GPU call 28 (1 GPU): the record of the final pipeline -- parity suite, bench lines of all BASELINE configs (C1 with
# both arms as the driver runs them, C2 / C3 / C4 with full-size image parity), ncu launch list + first-wave captures.
# The .ncu-rep files stay on the box except C1's first wave (a call brings back at most 64 MiB): raw / source pages as csv.gz.
set -u
O=gpurun_out
T=/tmp/ncu_final; mkdir -p $T
( timeout 2400 python -m pytest tests -m gpu -q > $O/r2z_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2z_gpu_tests.log ); tail -3 $O/r2z_gpu_tests.log
timeout 600 python bench.py > $O/r2z_bench_c1.json 2> $O/r2z_bench_c1.err; echo "c1 rc=$?"
timeout 600 python bench.py --impl reference > $O/r2z_bench_c1_ref.json 2> $O/r2z_bench_c1_ref.err; echo "c1 ref rc=$?"
timeout 900 python bench.py --workload materials --steps 3 --warmup 3 --parity-paths 3e8 > $O/r2z_bench_c2.json 2> $O/r2z_bench_c2.err; echo "c2 rc=$?"
timeout 900 python bench.py --workload ibl --steps 3 --warmup 3 --parity-paths 3e8 > $O/r2z_bench_c3.json 2> $O/r2z_bench_c3.err; echo "c3 rc=$?"
timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2z_bench_c4_spp16.json 2> $O/r2z_bench_c4_spp16.err; echo "c4 rc=$?"
timeout 900 python bench.py --workload intersect --steps 5 --warmup 3 > $O/r2z_bench_c5.json 2> $O/r2z_bench_c5.err; echo "c5 rc=$?"
timeout 900 python bench.py --workload cornell_spheres_bpt --steps 3 --warmup 3 > $O/r2z_bench_bpt.json 2> $O/r2z_bench_bpt.err; echo "bpt rc=$?"
for f in c1 c1_ref c2 c3 c4_spp16 c5 bpt; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2z_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],2), "cpu", (d.get("cpu_baseline") or {}).get("value"), "parity", (d.get("image_parity") or {}).get("ratio_to_floor"))
except Exception as e: print("$f", "ERR", e, open("$O/r2z_bench_$f.err").read()[-300:])
PY
done
NCU_L="ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
for W in cornell_spheres materials instanced; do
  X=""; [ $W = instanced ] && X="--spp 4"; [ $W = materials ] && X="--spp 16"
  timeout 900 $NCU_L --log-file $O/r2z_launches_$W.csv python tools/ncu_frame.py --workload $W $X > $O/r2z_ncu_$W.log 2>&1; echo "launch list $W rc=$?"
done
timeout 900 $NCU_L --log-file $O/r2z_launches_intersect.csv python tools/ncu_frame.py --workload intersect > $O/r2z_ncu_intersect.log 2>&1; echo "launch list intersect rc=$?"
K='materialKernel|extendKernel|surfaceKernel|shadowKernel|raygenKernel|emissionKernel'
FULL="ncu --profile-from-start off --set full --clock-control none --import-source on"
timeout 1200 $FULL -k regex:"$K" -c 8 -f -o $T/c1_w0 python tools/ncu_frame.py --workload cornell_spheres > $O/r2z_full_c1_w0.log 2>&1; echo "full c1 wave 0 rc=$?"
timeout 1200 $FULL -k regex:"$K" -s 16 -c 8 -f -o $T/c1_w2 python tools/ncu_frame.py --workload cornell_spheres > $O/r2z_full_c1_w2.log 2>&1; echo "full c1 wave 2 rc=$?"
timeout 1200 $FULL -k regex:'extendKernel|shadowKernel|surfaceKernel' -c 3 -f -o $T/c4_w0 python tools/ncu_frame.py --workload instanced --spp 4 > $O/r2z_full_c4.log 2>&1; echo "full c4 rc=$?"
timeout 1200 $FULL -k regex:'intersectBatchKernel' -c 1 -f -o $T/c5 python tools/ncu_frame.py --workload intersect > $O/r2z_full_c5.log 2>&1; echo "full c5 rc=$?"
for w in c1_w0 c1_w2 c4_w0 c5; do
  ncu -i $T/$w.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $O/r2z_${w}_raw.csv.gz
  ncu -i $T/$w.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/r2z_${w}_source.csv.gz
done
cp $T/c1_w0.ncu-rep $O/r2z_prof_c1_w0.ncu-rep
ls -la $O | head -60; du -sh $O
