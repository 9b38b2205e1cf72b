#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 1800 python -m pytest tests -q -m gpu --maxfail=6 > gpurun_out/t_gpu6.log 2>&1; echo "pytest gpu rc=$?"
tail -12 gpurun_out/t_gpu6.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench6_n1.log 2>&1; echo "bench n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench6_n2.log 2>&1; echo "bench n2 rc=$?"
tail -3 gpurun_out/bench6_n2.log | cut -c1-600
python - <<PY
import json
for f in ('gpurun_out/bench6_n1.log','gpurun_out/bench6_n2.log'):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]
        d=json.loads(l); print(f, d['n_gpus'], d['ms_per_step'], d['value']/1e9, d['stage_ms_per_step'], d['roofline']['frac'], d['config']['edges'], d['e2e'] and d['e2e']['value']/1e9)
    except Exception as e: print(f, 'ERR', e)
PY
