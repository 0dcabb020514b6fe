#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02i_gpu_tests.log | cut -c1-300
for wl in c3 c1 c2; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/r02i_bench_$wl.json 2> gpurun_out/r02i_bench_$wl.err; echo "$wl rc=$?"
done
JPCB_CHAIN_FFWD=0 timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r02i_bench_c3_nochain.json 2> gpurun_out/r02i_bench_c3_nochain.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02i_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'blocking', round(d['e2e_blocking_read']['value']), 'graph', d.get('e2e_graph') and round(d['e2e_graph']['value']), 'launches', d.get('gpu_launches'), d.get('launch_plans'))
    except Exception as e:
        print(f, 'FAILED', e); print(open(f.replace('.json','.err')).read()[-800:])
PY
JPCB_PLAN=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02i_launches_c3.csv python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02i_launches_c3.csv', errors='ignore')))
hdr=None; data=[]
for r in rows:
    if r and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr): data.append(dict(zip(hdr,r)))
print(len(data))
agg=collections.OrderedDict()
first=data[:len(data)]
for d in first[:60]:
    print(d['Kernel Name'][:60].ljust(62), d['Grid Size'].ljust(14), round(float(d['Metric Value'])/1e3,1),'us')
PY
