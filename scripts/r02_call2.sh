#!/bin/bash
# launch plans + host-side trimming: GPU tests, host profile of make_pc_step, benches of the launch-bound configs (plans on / off)
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02c_gpu_tests.log 2>&1; echo "tests rc=$?"; grep -E "plan cache|passed|failed|Error|error" gpurun_out/r02c_gpu_tests.log | tail -8
for wl in c3 c1; do timeout 300 python scripts/profile_e2e.py $wl > gpurun_out/r02c_profile_e2e_$wl.txt 2>&1; head -34 gpurun_out/r02c_profile_e2e_$wl.txt | cut -c1-150; done
for wl in c1 c2 c3 c4; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/r02c_bench_$wl.json 2> gpurun_out/r02c_bench_$wl.err; echo "$wl rc=$?"
  JPCB_PLAN=0 timeout 300 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/r02c_bench_${wl}_noplan.json 2> gpurun_out/r02c_bench_${wl}_noplan.err; echo "$wl noplan rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02c_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'graph', d.get('e2e_graph') and round(d['e2e_graph']['value']), 'launches', d.get('gpu_launches'))
    except Exception as e:
        print(f, 'FAILED', e); print(open(f.replace('.json','.err')).read()[-800:])
PY
