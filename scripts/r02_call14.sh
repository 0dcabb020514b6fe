#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02n_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02n_gpu_tests.log | cut -c1-200
run() { tag=$1; wl=$2; shift 2; env "$@" timeout 300 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/r02n_bench_${wl}_$tag.json 2> gpurun_out/r02n_bench_${wl}_$tag.err; }
run A c2 X=1
run B c2 JPCB_WARP8_UNROLL=1
run C c2 JPCB_WARP8_KMAX=1024
run D c2 JPCB_WARP8_KMAX=1024 JPCB_WARP8_UNROLL=1
run A c3 X=1
run B c3 JPCB_WARP8_UNROLL=1
run A c1 X=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02n_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
    except Exception as e:
        print(f, 'FAILED', e); print(open(f.replace('.json','.err')).read()[-600:])
PY
