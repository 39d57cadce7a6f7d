#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
for L in libslrgpu.so libslrgpu_nospread.so; do for T in 37888 75776 151552 303104; do
  SLR_BENCH_AB=1 SLRGPU_TAIL_PATHS=$T SLRGPU_LIB=$L timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$L', $T, round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'tail', d['config']['stage_ms_profiled_frame']['tailKernel'], 'waves', d['config']['waves_per_frame'], d['config']['tail_kernel'])"
done; done
SLR_BENCH_AB=1 SLRGPU_LIB=libslrgpu.so timeout 600 python bench.py --workload materials --spp 32 --steps 3 --warmup 3 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 spread', round(d['value'],1), d['config']['stage_ms_profiled_frame']['tailKernel'])"
SLR_BENCH_AB=1 SLRGPU_LIB=libslrgpu_nospread.so timeout 600 python bench.py --workload materials --spp 32 --steps 3 --warmup 3 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 nospread', round(d['value'],1), d['config']['stage_ms_profiled_frame']['tailKernel'])"
