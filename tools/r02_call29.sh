#!/bin/bash
# round 2, GPU call 29 (2 GPUs): the multi-GPU paths on the final pipeline -- slrgpu_render_multi tests, C1 weak and strong at N = 2
set -u
O=gpurun_out
run() {  # name N args...
  local name=$1 n=$2; shift 2
  if [ $n -eq 1 ]; then timeout 900 python bench.py --gpus 1 "$@" > $O/r2z_$name.json 2> $O/r2z_$name.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n "$@" > $O/r2z_$name.json 2> $O/r2z_$name.err; fi
  echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2z_$name.json").read().strip().splitlines()[-1])
    r=d["config"]["step_ms_per_rank"]
    print("$name", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), d["scaling"], "frame_spp", d["config"]["frame_spp"],
          "rank medians", [round(x["median"],2) for x in r], "rank max", [round(x["max"],2) for x in r])
except Exception as e: print("$name", "ERR", e, open("$O/r2z_$name.err").read()[-300:])
PY
}
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2
export SLR_BENCH_AB=1
run scale_c1_n1 1 --steps 10 --warmup 3
run scale_c1_n2_weak 2 --steps 10 --warmup 3
run scale_c1_n1_spp512 1 --steps 4 --warmup 3 --spp 512
run scale_c1_n2_strong 2 --steps 4 --warmup 3 --scaling strong --spp 512
unset SLR_BENCH_AB
# the bench lines again with the final captures' numbers in profiles/kernel_metrics.json / traffic.json
timeout 600 python bench.py > $O/r2z_bench_c1.json 2> $O/r2z_bench_c1.err; echo "c1 rc=$?"
timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2z_bench_c4_spp16.json 2> $O/r2z_bench_c4_spp16.err; echo "c4 rc=$?"
NCU_L="ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
timeout 600 $NCU_L --log-file $O/r2z_launches_intersect_grid500.csv python tools/ncu_frame.py --workload intersect --grid 500 > $O/r2z_ncu_intersect_grid500.log 2>&1; echo "launch list intersect grid 500 rc=$?"
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:intersectBatchKernel -c 1 -f -o /tmp/c5_500 python tools/ncu_frame.py --workload intersect --grid 500 > $O/r2z_full_c5_grid500.log 2>&1; ncu -i /tmp/c5_500.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $O/r2z_c5_grid500_raw.csv.gz
for f in c1 c4_spp16; do python - <<PY
import json
d=json.loads(open("$O/r2z_bench_$f.json").read().strip().splitlines()[-1])
print("$f", round(d["value"],2), d["unit"], "e2e", round(d["e2e"]["value"],2), "roofline", d["roofline"]["bound"], round(d["roofline"]["frac"],3), d["roofline"].get("lanes_active_of_32"))
PY
done
