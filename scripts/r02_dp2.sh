#!/bin/bash
# 2 GPUs: NCCL parity of the data-parallel step (functional API + PCTrainer), then the bench the driver runs (weak) and strong
P=29800
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P "$@"; P=$((P+1)); }
run scripts/dp_nccl_check.py > gpurun_out/r02_dp_nccl_parity_2gpu_final.txt 2>&1; echo "parity rc=$?"; grep -E "rel err|OK|Error|error|max" gpurun_out/r02_dp_nccl_parity_2gpu_final.txt | tail -14 | cut -c1-200
for mode in weak strong; do
  run bench.py --gpus 2 --steps 20 --warmup 5 --scaling $mode --no-cpu-baseline > gpurun_out/r02_final_bench_c4_${mode}_2gpu.json 2> gpurun_out/r02_final_bench_c4_${mode}_2gpu.err; echo "$mode rc=$?"
  python - gpurun_out/r02_final_bench_c4_${mode}_2gpu.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    print(sys.argv[1].split('/')[-1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'graph', d.get('e2e_graph') and round(d['e2e_graph']['value']))
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
run bench.py --gpus 2 --workload c3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_final_bench_c3_weak_2gpu.json 2> gpurun_out/r02_final_bench_c3_weak_2gpu.err; echo "c3 rc=$?"; tail -c 600 gpurun_out/r02_final_bench_c3_weak_2gpu.json | head -c 300; python -c "
import json;d=json.loads([l for l in open('gpurun_out/r02_final_bench_c3_weak_2gpu.json').read().strip().splitlines() if l.startswith('{')][-1]);print('c3 2gpu value',round(d['value']),'e2e',round(d['e2e']['value']),'graph',d.get('e2e_graph') and round(d['e2e_graph']['value']))"
