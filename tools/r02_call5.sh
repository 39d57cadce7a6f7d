#!/bin/bash
# round 2, GPU call 5 (2 GPUs): full parity suite (motion blur, two real devices behind slrgpu_render_multi), N = 1 / 2 bench lines
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -4
( timeout 2400 python -m pytest tests -m gpu -q > $O/r2e_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2e_gpu_tests.log )
tail -8 $O/r2e_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2e_bench_c1_n1.json 2> $O/r2e_bench_c1_n1.err; echo "bench n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2e_bench_c1_n2.json 2> $O/r2e_bench_c1_n2.err; echo "bench n2 weak rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --scaling strong --spp 512 > $O/r2e_bench_c1_n2_strong.json 2> $O/r2e_bench_c1_n2_strong.err; echo "bench n2 strong rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --spp 512 > $O/r2e_bench_c1_n1_spp512.json 2> $O/r2e_bench_c1_n1_spp512.err; echo "bench n1 spp512 rc=$?"
for f in c1_n1 c1_n2 c1_n2_strong c1_n1_spp512; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2e_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), d["scaling"], d["config"]["step_ms_per_rank"])
except Exception as e: print("$f", "ERR", e)
PY
done
