#!/bin/bash
# end-of-round records: full GPU test suite, smoke, every workload's bench line, the config-4 launch list
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_final_gpu_tests.log | cut -c1-300
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_final_smoke.log | cut -c1-200
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench_c4.json 2> gpurun_out/r02_final_bench_c4.err; echo "c4 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_bench_c4_reference_arm.json 2> gpurun_out/r02_final_ref.err; echo "ref rc=$?"
for wl in c1 c2 c3 c5; do timeout 300 python bench.py --workload $wl > gpurun_out/r02_final_bench_$wl.json 2> gpurun_out/r02_final_bench_$wl.err; echo "$wl rc=$?"; done
timeout 300 python bench.py --workload c5 --bpc-dtype bf16 --no-cpu-baseline > gpurun_out/r02_final_bench_c5_bf16.json 2> gpurun_out/r02_final_bench_c5_bf16.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_final_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', d.get('roofline') and round(d['roofline']['frac'],3), 'cpu', d.get('cpu_baseline') and round(d['cpu_baseline']['value'],1), 'launches', d.get('gpu_launches'))
    except Exception as e:
        print(f, 'FAILED', e); print(open(f.replace('.json','.err')).read()[-600:])
PY
JPCB_PLAN=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_final_launches_c4.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_final_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
