#!/bin/bash
# ncu --set full of two consecutive steady-state launches matching $1 (kernel regex), skipping $2 matches; output name $3
k=${1:-tc_staged}; s=${2:-100}; o=${3:-r02_staged}
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${o}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 2 -f -o gpurun_out/$o \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${o}_ncu.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/${o}_ncu.log | cut -c1-300
