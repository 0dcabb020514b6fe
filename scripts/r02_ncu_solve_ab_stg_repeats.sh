#!/bin/bash
m() { JPCB_PLAN=0 env "$@" timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,dram__bytes_read.sum --clock-control none -k regex:tc_staged -s 2 -c 1 --csv --log-file gpurun_out/r02t.csv python scripts/r02_infer_only.py > /dev/null 2>&1
echo "$*: $(grep tc_staged gpurun_out/r02t.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"; }
m X=warm
for rep in 1 2 3; do
m JPCB_TS_STG=0
m JPCB_TS_STG=1
m JPCB_TS_STG=1 JPCB_TS_HINTS=1
done
