#!/bin/bash
python scripts/r02_infer_only.py > gpurun_out/r02_infer_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_staged -s 2 -c 1 -f -o gpurun_out/r02_infer_solve python scripts/r02_infer_only.py > gpurun_out/r02_infer_ncu.log 2>&1
echo "ncu rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_plain_solve.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/r02_launches_c4_solve.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_solve_list.log 2>&1
echo "list rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err; tail -c 1500 gpurun_out/r02_bench_c4.json
