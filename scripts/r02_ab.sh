#!/bin/bash
# A/B of the staged-epilogue kernel on config 4 (JPCB_STAGED bit 0: GRAD, bit 1: ERR); parity tests first
set -x
python -m pytest tests -m gpu -x -q -k "bf16 or config4" > gpurun_out/r02_tests_bf16.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_bf16.log
tail -5 gpurun_out/r02_tests_bf16.log
for f in 0 1 2 3; do
  JPCB_STAGED=$f python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_ab_staged$f.json 2> gpurun_out/r02_ab_staged$f.err
  echo "staged=$f rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_ab_staged$f.json").read().strip().splitlines()[-1])
    print("staged=$f ms/step", d["ms_per_step"], "launch_ms", d["roofline"]["launch_ms"], "loss", d["loss"], d["clocks"])
except Exception as e:
    print("parse fail", e); print(open("gpurun_out/r02_ab_staged$f.err").read()[-2000:])
PY
done
