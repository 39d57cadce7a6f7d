#!/bin/bash
# round 2, GPU call 7 (8 GPUs): scaling evidence with the final pipeline -- C1 weak at 8, C1 strong (frame 512 spp) at 4 and 8,
# C4 (10 M instanced triangles, 1920x1080, frame 1024 spp = BASELINE configs[3]) at 8 and its one-GPU share (128 spp)
set -u
mkdir -p gpurun_out
O=gpurun_out
run() {  # name N args...
  local name=$1 n=$2; shift 2
  if [ $n -eq 1 ]; then timeout 900 python bench.py --gpus 1 "$@" > $O/r2g_$name.json 2> $O/r2g_$name.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n "$@" > $O/r2g_$name.json 2> $O/r2g_$name.err; fi
  echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2g_$name.json").read().strip().splitlines()[-1])
    r=d["config"]["step_ms_per_rank"]
    print("$name", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), d["scaling"], "frame_spp", d["config"]["frame_spp"],
          "rank medians", [round(x["median"],2) for x in r], "rank max", [round(x["max"],2) for x in r])
except Exception as e: print("$name", "ERR", e)
PY
}
nvidia-smi -L | wc -l
run c1_weak_n8 8 --steps 10 --warmup 3
run c1_strong_n8 8 --steps 10 --warmup 3 --scaling strong --spp 512
run c1_strong_n4 4 --steps 10 --warmup 3 --scaling strong --spp 512
run c4_strong_n8 8 --workload instanced --steps 3 --warmup 3 --scaling strong --spp 1024
run c4_n1_spp128 1 --workload instanced --steps 3 --warmup 3 --spp 128
