#!/bin/bash
# re-entry check: GPU tests on the restored tree, host-side profile of make_pc_step on the launch-bound configs, config 5 baselines
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02b_gpu_tests.log
for wl in c3 c1; do timeout 300 python scripts/profile_e2e.py $wl > gpurun_out/r02b_profile_e2e_$wl.txt 2>&1; head -40 gpurun_out/r02b_profile_e2e_$wl.txt | cut -c1-160; done
for wl in c1 c2 c3; do timeout 300 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/r02b_bench_$wl.json 2> gpurun_out/r02b_bench_$wl.err; echo "$wl rc=$?"; done
timeout 300 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r02b_bench_c5.json 2> gpurun_out/r02b_bench_c5.err; echo "c5 rc=$?"
timeout 300 python bench.py --workload c5 --bpc-dtype bf16 --no-cpu-baseline > gpurun_out/r02b_bench_c5_bf16.json 2> gpurun_out/r02b_bench_c5_bf16.err; echo "c5 bf16 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02b_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'graph', d.get('e2e_graph') and round(d['e2e_graph']['value']), 'launches', d.get('gpu_launches'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
