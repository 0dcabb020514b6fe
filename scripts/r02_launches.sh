#!/bin/bash
# per-launch device times (ncu, serialised) of one config-4 train step for JPCB_STAGED=$1
f=${1:-3}
export JPCB_STAGED=$f
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_plain_$f.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/r02_launches_staged$f.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_$f.log 2>&1
echo "rc=$?"
tail -3 gpurun_out/r02_ncu_$f.log
