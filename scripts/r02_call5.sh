#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02f_gpu_tests.log | cut -c1-300
for b in 512 1024 2048 4096; do timeout 120 python scripts/r02_b512.py $b; done 2>&1 | tee gpurun_out/r02f_b_local_times.txt
JPCB_PLAN=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02f_launches_c4_b512.csv python scripts/r02_b512.py 512 1 > gpurun_out/r02f_ncu_b512.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02f_launches_c4_b512.csv', errors='ignore')))
hdr=None; data=[]
for r in rows:
    if r and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr): data.append(r)
per=15
last=data[-2*per-0:-per] if len(data)>=2*per else data
for r in data[-30:]:
    d=dict(zip(hdr,r)); print(d['Kernel Name'][:50].ljust(52), d['Grid Size'].ljust(14), round(float(d['Metric Value'])/1e3,1),'us')
PY
for wl in c5; do timeout 300 python bench.py --workload c5 --no-cpu-baseline --bpc-dtype bf16 > gpurun_out/r02f_bench_c5_bf16.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r02f_bench_c5_bf16.json').read().strip().splitlines()[-1]);print('c5 bf16 ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']))"; done
