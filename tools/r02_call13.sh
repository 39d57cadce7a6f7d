#!/bin/bash
# final-pipeline bench lines of the other BASELINE configs, C2 / C3 with full-size image parity against two seeds of the reference
set -u
O=gpurun_out
timeout 900 python bench.py --workload materials --steps 3 --warmup 3 --parity-paths 3e8 > $O/r2x_bench_c2.json 2> $O/r2x_bench_c2.err
timeout 900 python bench.py --workload ibl --steps 3 --warmup 3 --parity-paths 3e8 > $O/r2x_bench_c3.json 2> $O/r2x_bench_c3.err
timeout 900 python bench.py --workload instanced --spp 16 --steps 3 --warmup 3 > $O/r2x_bench_c4_spp16.json 2> $O/r2x_bench_c4_spp16.err
timeout 900 python bench.py --workload intersect --steps 5 --warmup 3 > $O/r2x_bench_c5_default.json 2> $O/r2x_bench_c5_default.err
for f in c2 c3 c4_spp16 c5_default; do python - <<P
import json
try:
    d=json.loads(open('$O/r2x_bench_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1), 'cpu', d['cpu_baseline']['value'] if d.get('cpu_baseline') else None, 'parity', d.get('image_parity'))
except Exception as e:
    print('$f failed', e); print(open('$O/r2x_bench_$f.err').read()[-600:])
P
done
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -2
