#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r02_gpu_tests.log
bash scripts/r02_sweep.sh "JPCB_STAGED=7" "JPCB_STAGED=3"
