#!/bin/bash
m() { JPCB_PLAN=0 env "$@" timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:tc_staged -s 2 -c 1 --csv --log-file gpurun_out/r02q.csv python scripts/r02_infer_only.py > /dev/null 2>&1
echo "$*: $(grep tc_staged gpurun_out/r02q.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"; }
m JPCB_TS_UNIK=0
m JPCB_TS_UNIK=1 JPCB_TS_MMADELAY=0
m JPCB_TS_UNIK=1 JPCB_TS_MMADELAY=200
m JPCB_TS_UNIK=1 JPCB_TS_MMADELAY=500
m JPCB_TS_UNIK=1 JPCB_TS_MMADELAY=1000
m JPCB_TS_UNIK=1 JPCB_TS_MMADELAY=2000
m JPCB_TS_UNIK=0 JPCB_TS_MMADELAY=500
m JPCB_TS_UNIK=0 JPCB_TS_MMADELAY=1500
