#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chain or config3 or skips_scalings or fused" > gpurun_out/r02j_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02j_gpu_tests.log | cut -c1-300
for rep in 1 2; do
timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r02j_bench_c3_$rep.json 2> gpurun_out/r02j_bench_c3.err
JPCB_CHAIN_FFWD=0 timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r02j_bench_c3_nochain_$rep.json 2> gpurun_out/r02j_bench_c3_nochain.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02j_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'graph', d.get('e2e_graph') and round(d['e2e_graph']['value']), 'launches', d.get('gpu_launches'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
JPCB_PLAN=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:chain_ffwd -c 3 --csv --log-file gpurun_out/r02j_chain.csv python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; grep chain gpurun_out/r02j_chain.csv | cut -d, -f5,15 | tail -3
