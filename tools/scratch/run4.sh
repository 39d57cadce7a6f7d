#!/bin/bash
for v in "" _sb6 _sb8; do
  SLRGPU_LIB=libslrgpu$v.so timeout 200 python bench.py --steps 5 2>/dev/null > gpurun_out/run4$v.json
  echo "variant '$v'"; python tools/bench_summary.py gpurun_out/run4$v.json 2>/dev/null | grep -v "clocks\|roofline\|traversal"
done
