#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/multi_rank_check.py tiny c1_40x ref30k > gpurun_out/multi22_n$N.log 2>&1; echo "multi check rc=$?"; grep "MULTI_RANK" gpurun_out/multi22_n$N.log
for big in 1 0; do
PEDK_PART_BIG_TILE=$big timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$big bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench22_n${N}_big$big.log 2>&1; echo "bench big=$big rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench22_n?_big?.log')):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]
        d=json.loads(l); print(f, d['n_gpus'], round(d['ms_per_step'],1), round(d['value']/1e9,2), 'part_scatter', round(d['roofline']['all_kernels']['part_scatter']['ms_per_step'],2), 'exchange', round(d['stage_ms_per_step']['exchange'],1))
    except Exception as e: print(f, 'ERR', e); print(open(f).read()[-1500:])
PY
