#!/bin/bash
JPCB_PLAN=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_ffwd -s 3 -c 1 -o gpurun_out/r02_chain_ffwd python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_chain.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
