#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02k_gpu_tests.log | cut -c1-300
for wl in c3 c1; do timeout 300 python scripts/profile_e2e.py $wl > gpurun_out/r02k_profile_e2e_$wl.txt 2>&1; head -26 gpurun_out/r02k_profile_e2e_$wl.txt | cut -c1-150; done
for wl in c1 c3; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/r02k_bench_$wl.json 2> gpurun_out/r02k_bench_$wl.err; echo "$wl rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02k_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'blocking', round(d['e2e_blocking_read']['value']), 'graph', d.get('e2e_graph') and round(d['e2e_graph']['value']), d.get('launch_plans'))
    except Exception as e:
        print(f, 'FAILED', e); print(open(f.replace('.json','.err')).read()[-800:])
PY
