#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02o_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02o_gpu_tests.log | cut -c1-200; grep -n "^E  " gpurun_out/r02o_gpu_tests.log | head -6 | cut -c1-200
run() { tag=$1; wl=$2; shift 2; env "$@" timeout 300 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/r02o_bench_${wl}_$tag.json 2> gpurun_out/r02o_bench_${wl}_$tag.err; }
run auto c2 X=1
run off c2 JPCB_SIMT_KS=0
run auto c1 X=1
run force c1 JPCB_SIMT_KS=2
run auto c3 X=1
run force c3 JPCB_SIMT_KS=2
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02o_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'])
    except Exception as e:
        print(f, 'FAILED', e); print(open(f.replace('.json','.err')).read()[-600:])
PY
