#!/bin/bash
# sweep of the staged kernel's knobs on config 4: "ENV=VAL ..." strings, one bench run each
run() {
  tag=$(echo "$1" | tr ' =' '__')
  env $1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_sw_$tag.json 2> gpurun_out/r02_sw_$tag.err
  python - "$1" gpurun_out/r02_sw_$tag.json <<PY
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:40s} ms/step {d['ms_per_step']:.3f} launch_ms {d['roofline']['launch_ms']:.4f} loss {d['loss']:.7f} clk {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[2].replace('.json','.err')).read()[-1500:])
PY
}
for c in "$@"; do run "$c"; done
