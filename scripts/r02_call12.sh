#!/bin/bash
for wl in c2 c1; do
JPCB_PLAN=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02l_launches_$wl.csv python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "ncu $wl rc=$?"
done
python - <<'PY'
import csv,collections
for wl,per in (('c2',211),('c1',48)):
    rows=list(csv.reader(open(f'gpurun_out/r02l_launches_{wl}.csv', errors='ignore')))
    hdr=None; data=[]
    for r in rows:
        if r and r[0]=='ID': hdr=r; continue
        if hdr and len(r)==len(hdr): data.append(dict(zip(hdr,r)))
    step=data[:per]
    tot=sum(float(d['Metric Value']) for d in step)
    print(wl, len(data), 'launches captured; first step', per, 'launches', round(tot/1e3,1), 'us')
    agg=collections.OrderedDict()
    for d in step:
        k=(d['Kernel Name'][:44], d['Grid Size'])
        a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(d['Metric Value'])
    for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]:
        print(f"  {k[0]:46s} {k[1]:14s} n={c:4d} {t/1e3:8.1f} us avg {t/c/1e3:6.2f}")
PY
