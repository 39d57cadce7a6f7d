#!/bin/bash
# round 2, GPU call 21: ncu record of the current pipeline on C1 -- launch list of one frame, full captures of wave 0 and wave 2;
# the reports stay on the box (two of them exceed what a call may bring back), their raw / source pages come back as csv.gz
set -u
O=gpurun_out
T=/tmp/ncu_r2E; mkdir -p $T
NCU_L="ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
timeout 900 $NCU_L --log-file $O/r2E_launches_c1.csv python tools/ncu_frame.py --workload cornell_spheres > $O/r2E_ncu_c1.log 2>&1; echo "launch list c1 rc=$?"
K='materialKernel|extendKernel|surfaceKernel|shadowKernel|raygenKernel'
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$K" -c 7 -f -o $T/prof_c1_w0 python tools/ncu_frame.py --workload cornell_spheres > $O/r2E_ncu_full_c1_w0.log 2>&1; echo "full c1 wave 0 rc=$?"
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$K" -s 14 -c 7 -f -o $T/prof_c1_w2 python tools/ncu_frame.py --workload cornell_spheres > $O/r2E_ncu_full_c1_w2.log 2>&1; echo "full c1 wave 2 rc=$?"
for w in w0 w2; do
  ncu -i $T/prof_c1_$w.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $O/r2E_c1_${w}_raw.csv.gz
  ncu -i $T/prof_c1_$w.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/r2E_c1_${w}_source.csv.gz
done
cp $T/prof_c1_w2.ncu-rep $O/r2E_prof_c1_w2.ncu-rep
ls -la $O; du -sh $O
