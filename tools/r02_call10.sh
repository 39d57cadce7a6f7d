#!/bin/bash
# BPT kernel: resident blocks per SM (register cap) A/B
for L in libslrgpu.so libslrgpu_bpt3.so libslrgpu_bpt4.so libslrgpu_bpt6.so; do
  echo "== $L"
  BPT_DIRECT=1 SLRGPU_LIB=$L timeout 300 python tools/bpt_check.py 256 diffuse spheres materials instanced 2>&1 | grep " BPT "
done
