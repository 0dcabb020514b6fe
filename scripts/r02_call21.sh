#!/bin/bash
for rep in 1 2 3; do for h in 0 5; do
  JPCB_TS_HINTS=$h timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02u_bench_c4_h${h}_$rep.json 2> gpurun_out/r02u.err
  python -c "
import json;d=json.loads(open('gpurun_out/r02u_bench_c4_h${h}_$rep.json').read().strip().splitlines()[-1]);print('hints=$h rep=$rep ms/step',round(d['ms_per_step'],3),'solve launch ms',round(d['roofline']['launch_ms'],3),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']))"
done; done
