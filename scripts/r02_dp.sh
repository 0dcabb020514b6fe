#!/bin/bash
# data-parallel checks on N GPUs of one box: NCCL parity, then config 4 weak and strong scaling (fp32 and bf16 gradient payload)
N=${1:-2}
P=29500
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P "$@"; P=$((P+1)); }
run scripts/dp_nccl_check.py > gpurun_out/r02_dp_nccl_parity_${N}gpu.txt 2>&1; echo "parity rc=$?"; grep -E "rel err|OK|Error|error" gpurun_out/r02_dp_nccl_parity_${N}gpu.txt | tail -12
for mode in weak strong; do
  for gd in f32 bf16; do
    run bench.py --gpus $N --steps 20 --warmup 5 --scaling $mode --grad-dtype $gd --no-cpu-baseline > gpurun_out/r02_bench_c4_${mode}_${gd}_${N}gpu.json 2> gpurun_out/r02_bench_c4_${mode}_${gd}_${N}gpu.err
    python - gpurun_out/r02_bench_c4_${mode}_${gd}_${N}gpu.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['clocks'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
  done
done
