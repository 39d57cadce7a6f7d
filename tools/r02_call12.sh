#!/bin/bash
for L in libslrgpu.so libslrgpu_bpt3.so libslrgpu_bpt4.so libslrgpu_bpt5.so; do
  echo "== $L"
  BPT_DIRECT=1 SLRGPU_LIB=$L timeout 300 python tools/bpt_check.py 64 spheres materials instanced 2>&1 | grep " BPT " | cut -c1-40,150-260
  SLRGPU_LIB=$L python tools/bpt_profile.py spheres 512 16 2>&1 | tail -1
done
