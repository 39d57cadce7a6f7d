#!/bin/bash
# round 2, GPU call 1: parity tests, C1 bench (both arms), single-frame launch lists (traffic) for C1-C5, full captures for C4 / C5
set -u
mkdir -p gpurun_out
O=gpurun_out
NCU_L="ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
( timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2a_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2a_gpu_tests.log ) 
tail -3 $O/r2a_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2a_bench_c1.json 2> $O/r2a_bench_c1.err; echo "bench c1 rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2a_bench_c1_ref.json 2> $O/r2a_bench_c1_ref.err; echo "bench ref rc=$?"
for W in cornell_spheres:64 materials:16 ibl:16; do
  N=${W%%:*}; S=${W##*:}
  timeout 900 python tools/ncu_frame.py --workload $N --spp $S > $O/r2a_frame_$N.json 2> $O/r2a_frame_$N.err && \
  timeout 900 $NCU_L --log-file $O/r2a_launches_$N.csv python tools/ncu_frame.py --workload $N --spp $S > $O/r2a_ncu_$N.log 2>&1
  echo "launch list $N rc=$?"
done
# C4: 10 M instanced triangles (host build ~15 s per run)
timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2a_bench_c4_spp16.json 2> $O/r2a_bench_c4.err; echo "bench c4 rc=$?"
timeout 900 $NCU_L --log-file $O/r2a_launches_instanced.csv python tools/ncu_frame.py --workload instanced --spp 4 > $O/r2a_ncu_instanced.log 2>&1; echo "launch list c4 rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'extendKernel|shadowKernel' -c 2 -f -o $O/r2a_prof_c4 python tools/ncu_frame.py --workload instanced --spp 4 > $O/r2a_ncu_full_c4.log 2>&1; echo "full c4 rc=$?"
# C5: 10 M-triangle flat mesh
timeout 900 $NCU_L --log-file $O/r2a_launches_intersect.csv python tools/ncu_frame.py --workload intersect --grid 2236 --rays 16777216 > $O/r2a_ncu_intersect.log 2>&1; echo "launch list c5 rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:intersectBatchKernel -c 1 -f -o $O/r2a_prof_c5 python tools/ncu_frame.py --workload intersect --grid 2236 --rays 16777216 > $O/r2a_ncu_full_c5.log 2>&1; echo "full c5 rc=$?"
# C1 first-wave full capture of every kernel family (baseline for the kernel work)
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'materialKernel|extendKernel|surfaceKernel|shadowKernel|raygenKernel' -c 7 -f -o $O/r2a_prof_c1 python tools/ncu_frame.py --workload cornell_spheres > $O/r2a_ncu_full_c1.log 2>&1; echo "full c1 rc=$?"
cut -c1-400 $O/r2a_bench_c1.json
