#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02w_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02w_gpu_tests.log | cut -c1-200; grep -n "^E  " gpurun_out/r02w_gpu_tests.log | head -5 | cut -c1-200
timeout 300 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r02w_bench_c5.json 2>/dev/null
timeout 300 python bench.py --workload c5 --bpc-dtype bf16 --no-cpu-baseline > gpurun_out/r02w_bench_c5_bf16.json 2>/dev/null
python - <<'PY'
import json
for f in ('gpurun_out/r02w_bench_c5.json','gpurun_out/r02w_bench_c5_bf16.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d.get('launch_plans'))
PY
