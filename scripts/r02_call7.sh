#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02h_gpu_tests.log | cut -c1-300
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --workload c5 --no-cpu-baseline $EXTRA > gpurun_out/r02h_bench_c5_$tag.json 2> gpurun_out/r02h_bench_c5_$tag.err; echo "c5 $tag rc=$?"; }
EXTRA="" run x3 JPCB_PLAN=16
EXTRA="" run x3_noplan JPCB_PLAN=0
EXTRA="--bpc-dtype bf16" run bf16 JPCB_PLAN=16
EXTRA="--bpc-dtype bf16" run bf16_noplan JPCB_PLAN=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02h_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'launches', d.get('gpu_launches'))
    except Exception as e:
        print(f, 'FAILED', e); print(open(f.replace('.json','.err')).read()[-800:])
PY
