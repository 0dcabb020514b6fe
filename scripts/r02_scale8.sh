#!/bin/bash
# one 8-GPU box: config 4 strong scaling at 2 / 4 / 8 ranks (fp32 gradient payload), bf16 payload at 8, weak scaling at 8
P=29600
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P "$@"; P=$((P+1)); }
job() { n=$1; mode=$2; gd=$3; out=gpurun_out/r02_bench_c4_${mode}_${gd}_${n}gpu
  run $n bench.py --gpus $n --steps 20 --warmup 5 --scaling $mode --grad-dtype $gd --no-cpu-baseline > $out.json 2> $out.err
  python - $out.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    print(sys.argv[1].split('/')[-1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'graph', d.get('e2e_graph') and round(d['e2e_graph']['value']), d['clocks'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c4_1gpu_samebox.json 2> gpurun_out/r02_bench_c4_1gpu_samebox.err
python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_c4_1gpu_samebox.json').read().strip().splitlines()[-1]);print('1gpu value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']))"
job 8 strong f32
job 8 strong bf16
job 4 strong f32
job 2 strong f32
job 8 weak f32
