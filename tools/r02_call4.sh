#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_sbvh_traversal.py tests/test_gpu_intersect.py -m gpu -q > $O/r2d_sbvh_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2d_sbvh_tests.log ); tail -3 $O/r2d_sbvh_tests.log
timeout 900 python tools/diag_blocks.py Cornell_Box_ColorChecker.txt 96 512 > $O/r2d_diag_cc.log 2>&1; echo "diag rc=$?"; tail -22 $O/r2d_diag_cc.log
timeout 900 python tools/diag_blocks.py diffuse 96 512 > $O/r2d_diag_diffuse.log 2>&1; tail -9 $O/r2d_diag_diffuse.log
