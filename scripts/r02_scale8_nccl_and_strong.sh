#!/bin/bash
P=29700
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P "$@"; P=$((P+1)); }
run 8 scripts/r02_nccl_micro.py 2>&1 | grep -E "GPUs|round trip" | tee gpurun_out/r02_nccl_allreduce_8gpu.txt
job() { n=$1; mode=$2; gd=$3; tag=$4; shift 4; out=gpurun_out/r02_bench_c4_${mode}_${gd}_${n}gpu_$tag
  run $n bench.py --gpus $n --steps 20 --warmup 5 --scaling $mode --grad-dtype $gd --no-cpu-baseline "$@" > $out.json 2> $out.err
  python - $out.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    print(sys.argv[1].split('/')[-1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'graph', d.get('e2e_graph') and round(d['e2e_graph']['value']), d['clocks'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
job 8 strong f32 auto
job 8 strong bf16 auto
