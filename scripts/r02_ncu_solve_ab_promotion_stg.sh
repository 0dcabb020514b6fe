#!/bin/bash
m() { JPCB_PLAN=0 env "$@" timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tc_staged -s 2 -c 1 --csv --log-file gpurun_out/r02r.csv python scripts/r02_infer_only.py > /dev/null 2>&1
echo "$*: $(grep tc_staged gpurun_out/r02r.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"; }
m X=0
m JPCB_TS_STG=1
m JPCB_TS_PROMO=128
m JPCB_TS_PROMO=64
m JPCB_TS_PROMO=0
m JPCB_TS_PROMO=128 JPCB_TS_STG=1
m X=0
