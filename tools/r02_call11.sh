#!/bin/bash
# BPT: bench line at C1's scene, ncu of the two stage kernels, PT headline re-check after the bsdf.cuh / camera.cuh refactor
set -u
O=gpurun_out
timeout 900 python bench.py --workload cornell_spheres_bpt --steps 5 --warmup 3 > $O/r2r_bench_bpt.json 2> $O/r2r_bench_bpt.err
tail -c 1500 $O/r2r_bench_bpt.json
SLR_BENCH_AB=1 timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PT C1', round(d['value'],1), 'ms', round(d['ms_per_step'],2), d['config']['stage_ms_profiled_frame'])"
python tools/bpt_profile.py spheres 256 16 > $O/r2r_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'subpathKernel|connectKernel' -s 2 -c 2 -o $O/r2r_bpt_stages python tools/bpt_profile.py spheres 256 16 > $O/r2r_ncu.log 2>&1
tail -2 $O/r2r_plain.log
