#!/bin/bash
N=${1:-2}
for wl in c4 c3; do
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --workload $wl --no-cpu-baseline > gpurun_out/r02_bench_${wl}_weak_${N}gpu.json 2> gpurun_out/r02_bench_${wl}_weak_${N}gpu.err; echo "bench $wl rc=$?"; tail -1 gpurun_out/r02_bench_${wl}_weak_${N}gpu.err
python - gpurun_out/r02_bench_${wl}_weak_${N}gpu.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'e2e_graph', d['e2e_graph'] and round(d['e2e_graph']['value']))
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open(sys.argv[1].replace('.json','.err')).read()[-2500:])
PY
done
