#!/bin/bash
for v in "" _tb10 _tb12; do
  SLRGPU_LIB=libslrgpu$v.so timeout 200 python bench.py --steps 5 2>/dev/null > gpurun_out/var$v.json
  echo "variant '$v'"; python tools/bench_summary.py gpurun_out/var$v.json | grep -v "clocks\|roofline\|e2e:"
done
for pool in 1048576 4194304 8388608; do
  timeout 200 python bench.py --steps 5 --pool $pool 2>/dev/null > gpurun_out/var_pool$pool.json
  echo "pool $pool"; python tools/bench_summary.py gpurun_out/var_pool$pool.json | head -1
done
