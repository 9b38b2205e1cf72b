#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_kernels.py -q -m gpu > gpurun_out/t_gpu7.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_gpu7.log
PEDK_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench7_n2_trace.log 2>&1; echo "trace rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench7_n2.log 2>&1; echo "bench n2 rc=$?"
python - <<PY
import json
for f in ('gpurun_out/bench7_n2.log',):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]
        d=json.loads(l); print(f, d['n_gpus'], d['ms_per_step'], d['value']/1e9, d['stage_ms_per_step'], d['config']['edges'], d['e2e'] and d['e2e']['value']/1e9)
    except Exception as e: print(f, 'ERR', e)
PY
