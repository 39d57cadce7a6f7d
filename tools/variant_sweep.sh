#!/bin/bash
# A/B of tuning variants built side by side by slr_b200/csrc/Makefile (VARIANT=_name EXTRA=-D...): device frame time of C1
# (and C2 at 32 spp for the shade-kernel variants), intersect bench for the walk variants. Run under gpurun.
for L in libslrgpu.so $(cd slr_b200/lib && ls libslrgpu_*.so 2>/dev/null); do
  echo "== $L"
  SLRGPU_LIB=$L python tools/tail_sweep.py cornell_spheres 5 --tails 37888 2>/dev/null | tail -1 | cut -c1-330
  case $L in *_s*i*|libslrgpu.so) SLRGPU_LIB=$L python bench.py --workload intersect --steps 5 --warmup 3 --cpu-sample 20000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('intersect', round(d['value'],1), d['unit'])";; esac
  case $L in *_mb*|*_sb*|libslrgpu.so) SLRGPU_LIB=$L python tools/tail_sweep.py materials 2 --tails 37888 --spp 32 2>/dev/null | tail -1 | cut -c1-330;; esac
done
