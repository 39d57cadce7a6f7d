#!/bin/bash
# round 2, GPU call 19: C4 check of the walk as committed, then the ncu record of the current pipeline on C1 (host-driven loop:
# the conditional graph node of the device-driven loop cannot be profiled kernel by kernel)
set -u
O=gpurun_out
export SLR_BENCH_AB=1
SLRGPU_LIB=libslrgpu.so timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2D_c4.json 2> $O/r2D_c4.err
python - <<PY
import json
d=json.loads(open("$O/r2D_c4.json").read().strip().splitlines()[-1])
print("c4", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), d["config"].get("stage_ms_profiled_frame"))
PY
unset SLR_BENCH_AB
NCU_L="ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
timeout 900 $NCU_L --log-file $O/r2D_launches_c1.csv python tools/ncu_frame.py --workload cornell_spheres > $O/r2D_ncu_c1.log 2>&1; echo "launch list c1 rc=$?"
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'materialKernel|extendKernel|surfaceKernel|shadowKernel|raygenKernel' -c 21 -f -o $O/r2D_prof_c1 python tools/ncu_frame.py --workload cornell_spheres > $O/r2D_ncu_full_c1.log 2>&1; echo "full c1 rc=$?"
ls -la $O/r2D_prof_c1.ncu-rep; wc -l $O/r2D_launches_c1.csv
