#!/bin/bash
for v in "" _mb4 _mb5 _w8 _w2; do
  SLRGPU_LIB=libslrgpu$v.so timeout 200 python bench.py --steps 5 2>/dev/null > gpurun_out/var$v.json
  echo "variant '$v'"; python tools/bench_summary.py gpurun_out/var$v.json | head -1
done
for v in "" _w8 _w2; do
  echo "intersect variant '$v'"; SLRGPU_LIB=libslrgpu$v.so timeout 200 python bench.py --workload intersect --steps 5 2>/dev/null | python tools/bench_summary.py | head -1
done
