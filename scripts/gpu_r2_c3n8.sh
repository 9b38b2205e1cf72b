#!/bin/bash
# gpurun --gpus 8: the largest config (configs[3]) on 8 ranks
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 5 --warmup 3 --workload configs3 --no-cpu-baseline --no-e2e > gpurun_out/r2u_bench_configs3_n$N.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/r2u_bench_configs3_n$N.log') if x.startswith('{')][-1]
d=json.loads(l); print(d['n_gpus'], round(d['ms_per_step'],1), round(d['value']/1e9,2), {k:round(v,1) for k,v in d['stage_ms_per_step'].items()}, 'digests', d['parity'] and d['parity'].get('edges_sha','')[:16], d['parity'] and d['parity']['table_sum'])
PY
