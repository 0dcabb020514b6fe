#!/bin/bash
# L2 eviction hints of the staged kernel: ncu cycles / DRAM bytes of the dense solve launch, 3 processes per setting (placement noise)
m() { JPCB_PLAN=0 env "$@" timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,dram__bytes_read.sum --clock-control none -k regex:tc_staged -s 2 -c 1 --csv --log-file gpurun_out/r02s.csv python scripts/r02_infer_only.py > /dev/null 2>&1
echo "$*: $(grep tc_staged gpurun_out/r02s.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"; }
for rep in 1 2 3; do
for h in 0 1 7 5; do m JPCB_TS_HINTS=$h; done
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "whole_solve" > gpurun_out/r02s_tests.log 2>&1; echo "tests rc=$?"
JPCB_TS_HINTS=7 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "whole_solve or config4" > gpurun_out/r02s_tests7.log 2>&1; echo "tests(hints=7) rc=$?"; tail -1 gpurun_out/r02s_tests7.log
