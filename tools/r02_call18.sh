#!/bin/bash
# round 2, GPU call 18: the walk with the record loads issued together (instanced kernels) -- parity subset, bench lines, then
# the ncu record of the current pipeline on C1: launch list of one frame and full captures of the first three waves
set -u
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "intersect or sbvh or occlu or traversal or nested or probe or motion" 2>&1 | tail -3
export SLR_BENCH_AB=1
for L in libslrgpu.so; do
  SLRGPU_LIB=$L timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2C_c1_$L.json 2> $O/r2C_c1_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2C_c4_$L.json 2> $O/r2C_c4_$L.err
  SLRGPU_LIB=$L timeout 900 python bench.py --workload intersect --steps 5 --warmup 3 --cpu-sample 20000 > $O/r2C_c5_$L.json 2> $O/r2C_c5_$L.err
  for W in c1 c4 c5; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2C_${W}_$L.json").read().strip().splitlines()[-1])
    print("$W $L", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), d["config"].get("stage_ms_profiled_frame"))
except Exception as e: print("$W $L", "ERR", e, open("$O/r2C_${W}_$L.err").read()[-400:])
PY
  done
done
unset SLR_BENCH_AB
NCU_L="ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
timeout 900 $NCU_L --log-file $O/r2C_launches_c1.csv python tools/ncu_frame.py --workload cornell_spheres > $O/r2C_ncu_c1.log 2>&1; echo "launch list c1 rc=$?"
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'materialKernel|extendKernel|surfaceKernel|shadowKernel|raygenKernel' -c 21 -f -o $O/r2C_prof_c1 python tools/ncu_frame.py --workload cornell_spheres > $O/r2C_ncu_full_c1.log 2>&1; echo "full c1 rc=$?"
ls -la $O/r2C_prof_c1.ncu-rep
