#!/bin/bash
# end of round, one 8-GPU box: the driver's weak-scaling line at 8 GPUs and the strong-scaling line (4096 rows split 8 ways)
P=29900
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P "$@"; P=$((P+1)); }
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_final_bench_c4_1gpu_of8.json 2>/dev/null
python -c "
import json;d=json.loads(open('gpurun_out/r02_final_bench_c4_1gpu_of8.json').read().strip().splitlines()[-1]);print('1gpu value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']))"
for mode in weak strong; do
  run 8 bench.py --gpus 8 --steps 20 --warmup 5 --scaling $mode --no-cpu-baseline > gpurun_out/r02_final_bench_c4_${mode}_8gpu.json 2> gpurun_out/r02_final_bench_c4_${mode}_8gpu.err
  python - gpurun_out/r02_final_bench_c4_${mode}_8gpu.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    print(sys.argv[1].split('/')[-1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'graph', d.get('e2e_graph') and round(d['e2e_graph']['value']), d['clocks'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
