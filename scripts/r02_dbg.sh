#!/bin/bash
# per-launch time + cycles of the staged kernel with parts of it switched off (JPCB_TS_DBG; results are garbage)
for d in "$@"; do
  export JPCB_TS_DBG=$d
  ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max --clock-control none -k regex:tc_staged -s 100 -c 8 --csv --log-file gpurun_out/r02_dbg_$d.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_dbg_$d.log 2>&1
  echo "dbg=$d rc=$?"
  python - gpurun_out/r02_dbg_$d.csv <<'PY'
import csv,sys
vals={}
for r in csv.reader(open(sys.argv[1])):
    if len(r)>10 and r[0].isdigit():
        vals.setdefault(int(r[0]),{})[r[-3]]=float(r[-1].replace(',',''))
for k in sorted(vals): print('   launch',k,{a:round(b,1) for a,b in vals[k].items()})
PY
done
