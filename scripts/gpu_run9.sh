#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --maxfail=6 > gpurun_out/t_gpu9.log 2>&1; echo "pytest gpu rc=$?"
tail -6 gpurun_out/t_gpu9.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench9_n1.log 2>&1; echo "bench n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench9_n2.log 2>&1; echo "bench n2 rc=$?"
PEDK_P2P=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench9_n2_nccl.log 2>&1; echo "bench n2 nccl rc=$?"
PEDK_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench9_n2_trace.log 2>&1; echo "trace rc=$?"
tail -3 gpurun_out/bench9_n2.log | cut -c1-400
python - <<PY
import json
for f in ('gpurun_out/bench9_n1.log','gpurun_out/bench9_n2.log','gpurun_out/bench9_n2_nccl.log'):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]
        d=json.loads(l); print(f, d['n_gpus'], round(d['ms_per_step'],1), round(d['value']/1e9,2), {k:round(v,1) for k,v in d['stage_ms_per_step'].items()}, round(d['roofline']['frac'],3), d['config']['edges'], d['e2e'] and round(d['e2e']['value']/1e9,2))
    except Exception as e: print(f, 'ERR', e)
PY
