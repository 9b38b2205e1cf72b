#!/bin/bash
# gpurun --gpus N: sharded path against the oracle (incl. the overflow-and-redo case), bench at N with parity,
# a PEDK_TRACE step, optional extra workload (WL=configs3 ...)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
TAG=${TAG:-r2m}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/multi_rank_check.py tiny adversarial clip50 c1_40x ref30k lowcomplexity > gpurun_out/${TAG}_check_n$N.log 2>&1; echo "multi check rc=$?"; grep -v "^\*\|^$\|OMP_NUM" gpurun_out/${TAG}_check_n$N.log | tail -24
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n$N.log 2>&1; echo "bench rc=$?"
PEDK_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus $N --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_trace_n$N.log 2>&1; echo "trace rc=$?"
if [ -n "$WL" ]; then
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline --workload $WL > gpurun_out/${TAG}_bench_${WL}_n$N.log 2>&1; echo "bench $WL rc=$?"
fi
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench*_n$N.log')):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]
        d=json.loads(l); print(f, d['n_gpus'], round(d['ms_per_step'],1), round(d['value']/1e9,2), {k:round(v,1) for k,v in d['stage_ms_per_step'].items()}, d['config']['edges'], 'e2e', d['e2e'] and round(d['e2e']['value']/1e9,2), 'parity', d['parity'] and d['parity']['match'])
        for k,v in d['roofline']['all_kernels'].items(): print('   ',k, round(v['ms_per_step'],2), round(v['frac'],3))
    except Exception as e: print(f, 'ERR', e); print(open(f).read()[-2500:])
PY
