#!/bin/bash
# cycles of the whole-solve launch (dense, config 4) under ncu with / without the MMA issuer's tile decode
for v in 1 0 1 0; do
JPCB_PLAN=0 JPCB_TS_UNIK=$v timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:tc_staged -s 2 -c 1 --csv --log-file gpurun_out/r02p_solve_unik$v.csv python scripts/r02_infer_only.py > /dev/null 2>&1
echo "unik=$v: $(grep tc_staged gpurun_out/r02p_solve_unik$v.csv | awk -F'","' '{print $(NF-2), $NF}' | tr '\n' ' ')"
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "whole_solve or config4" > gpurun_out/r02p_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02p_tests.log
