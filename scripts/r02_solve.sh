#!/bin/bash
# whole-solve launch: parity tests, then bench A/B (JPCB_STAGED=7 whole-solve, 3 per-phase staged, 0 round-1 kernels)
timeout 600 python -m pytest tests -m gpu -x -q -k "whole_solve or bf16 or config4" > gpurun_out/r02_solve_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r02_solve_tests.log
bash scripts/r02_sweep.sh "JPCB_STAGED=7" "JPCB_STAGED=3" "JPCB_STAGED=7" "JPCB_STAGED=0"
