#!/bin/bash
# Per-kernel durations of the two inference phases of the c4 bench (ERR plain, GRAD main, GRAD layer-1) under ncu.
# usage (on the GPU box, via gpurun): bash scripts/ncu_phase_times.sh <tag>
set -e
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/pt_plain_$1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "regex:tc_grouped_gemm<\(int\)2, \(int\)[12]," -s 60 -c 60 --csv \
    --log-file gpurun_out/pt_$1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/pt_ncu_$1.log 2>&1
python - "$1" <<'PY'
import csv, sys, collections
rows = [l for l in open(f"gpurun_out/pt_{sys.argv[1]}.csv") if not l.startswith("==")]
agg = collections.defaultdict(list)
for r in csv.DictReader(rows):
    agg[r["Kernel Name"].split("tc_grouped_gemm")[1][:60]].append(float(r["Metric Value"].replace(",", "")) / 1e3)
for k, v in agg.items():
    v.sort()
    print(f"{sys.argv[1]} {k}: n={len(v)} median={v[len(v)//2]:.1f} us max={v[-1]:.1f} us")
PY
