#!/bin/bash
# gpurun --gpus N: bench.py at N (strong scaling line with parity), nothing else
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); TAG=${TAG:-r2s}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n$N.log 2>&1; echo "bench rc=$?"
if [ "$N" = "2" ]; then
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/${TAG}_tests_multi.log 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/${TAG}_tests_multi.log
PEDK_GUARD=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tests/multi_rank_check.py tiny adversarial c1_40x ref30k lowcomplexity ref30k_cnv > gpurun_out/${TAG}_check_guard_n2.log 2>&1; echo "guard multi check rc=$?"; grep "MULTI_RANK\|PEDK_GUARD" gpurun_out/${TAG}_check_guard_n2.log | head
fi
python - <<PY
import json
l=[x for x in open('gpurun_out/${TAG}_bench_n$N.log') if x.startswith('{')][-1]
d=json.loads(l); print(d['n_gpus'], round(d['ms_per_step'],1), round(d['value']/1e9,2), {k:round(v,1) for k,v in d['stage_ms_per_step'].items()}, 'e2e', d['e2e'] and round(d['e2e']['value']/1e9,2), 'parity', d['parity'] and d['parity']['match'])
print({k:round(v['ms_per_step'],2) for k,v in d['roofline']['all_kernels'].items()})
PY
