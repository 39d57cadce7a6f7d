#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_intersect.py -m gpu -x -q 2>&1 | tail -2
for pool in 0 8388608 16777216; do
  timeout 200 python bench.py --steps 5 --pool $pool 2>/dev/null > gpurun_out/run3_pool$pool.json
  echo "pool $pool"; python tools/bench_summary.py gpurun_out/run3_pool$pool.json 2>/dev/null | grep -v "clocks\|roofline\|e2e:"
done
timeout 300 python bench.py --workload intersect --steps 5 2>/dev/null | python tools/bench_summary.py 2>/dev/null | head -1
timeout 600 python tools/render_check.py instanced_full --width 1920 --height 1080 --spp 16 --no-ref --out gpurun_out/render_check_full 2>&1 | grep "^{" | cut -c250-520
