#!/bin/bash
# gpurun --gpus 2: peer-store microbenchmark, then the TMA bulk-store variant (PEDK_PART_BULK=1) of the two
# scatter kernels: parity subset + multi-rank check + bench A/B at N=1 and N=2
mkdir -p gpurun_out
timeout 600 scripts/bin/microbench_peer_stores > gpurun_out/r2d_microbench_peer_stores.txt 2>&1; echo "microbench rc=$?"; tail -5 gpurun_out/r2d_microbench_peer_stores.txt
PEDK_PART_BULK=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny or adversarial or dup or other_k or overflow or many_instances or corner or c1_40x or mixed or clip or medium" > gpurun_out/r2d_tests_bulk.log 2>&1; echo "bulk parity rc=$?"; tail -3 gpurun_out/r2d_tests_bulk.log
PEDK_PART_BULK=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/multi_rank_check.py tiny adversarial c1_40x ref30k lowcomplexity > gpurun_out/r2d_check_bulk_n2.log 2>&1; echo "bulk multi check rc=$?"; grep "MULTI_RANK" gpurun_out/r2d_check_bulk_n2.log
PEDK_GUARD=1 timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2d_tests_guard.log 2>&1; echo "guard-zone test run rc=$?"; tail -3 gpurun_out/r2d_tests_guard.log
PEDK_GUARD=1 PEDK_PART_BULK=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tests/multi_rank_check.py tiny c1_40x lowcomplexity > gpurun_out/r2d_check_guard_n2.log 2>&1; echo "guard multi check rc=$?"; grep "MULTI_RANK\|PEDK_GUARD" gpurun_out/r2d_check_guard_n2.log | head
for bulk in 0 1; do
PEDK_PART_BULK=$bulk timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2d_bench_n1_bulk$bulk.log 2>&1; echo "bench n1 bulk=$bulk rc=$?"
PEDK_PART_BULK=$bulk PEDK_PART_BIG_TILE=2 timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2d_bench_n1big_bulk$bulk.log 2>&1; echo "bench n1 big bulk=$bulk rc=$?"
PEDK_PART_BULK=$bulk timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$bulk bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2d_bench_n2_bulk$bulk.log 2>&1; echo "bench n2 bulk=$bulk rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2d_bench_*.log')):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]
        d=json.loads(l); k=d['roofline']['all_kernels']; print(f, d['n_gpus'], round(d['ms_per_step'],1), 'part_scatter', round(k['part_scatter']['ms_per_step'],2), 'join_scatter', round(k['join_scatter']['ms_per_step'],2), 'exchange', round(d['stage_ms_per_step']['exchange'],1), 'parity', d['parity'] and d['parity']['match'])
    except Exception as e: print(f, 'ERR', e); print(open(f).read()[-1500:])
PY
