#!/bin/bash
# C4 on 8 GPUs (and 1 GPU), C2/C3 on 1 GPU
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --workload instanced --steps 3 --warmup 3 > gpurun_out/c4_n8.json 2> gpurun_out/c4_err.log; echo "c4 n8 rc=$?"
timeout 600 python bench.py --workload instanced --steps 3 --warmup 3 > gpurun_out/c4_n1.json 2>> gpurun_out/c4_err.log; echo "c4 n1 rc=$?"
timeout 600 python bench.py --workload materials --steps 3 --warmup 3 > gpurun_out/c2_n1.json 2>> gpurun_out/c4_err.log; echo "c2 rc=$?"
timeout 600 python bench.py --workload ibl --steps 3 --warmup 3 > gpurun_out/c3_n1.json 2>> gpurun_out/c4_err.log; echo "c3 rc=$?"
python tools/bench_summary.py gpurun_out/c4_n8.json gpurun_out/c4_n1.json gpurun_out/c2_n1.json gpurun_out/c3_n1.json 2>/dev/null | grep -v "clocks"
grep -i "error\|traceback" -A3 gpurun_out/c4_err.log | head -20
