#!/bin/bash
# LSU store path of the staged kernel: tests, then A/B on the config-4 bench (same box, alternating)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02g_gpu_tests.log | cut -c1-300
for rep in 1 2; do for stg in 0 1; do
  JPCB_TS_STG=$stg timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02g_bench_c4_stg${stg}_$rep.json 2> gpurun_out/r02g_bench_c4_stg${stg}_$rep.err
  python -c "
import json;d=json.loads(open('gpurun_out/r02g_bench_c4_stg${stg}_$rep.json').read().strip().splitlines()[-1]);print('stg=$stg rep=$rep ms/step',round(d['ms_per_step'],3),'solve launch ms',round(d['roofline']['launch_ms'],3),'e2e',round(d['e2e']['value']),d['clocks'])"
done; done
