#!/bin/bash
# 1 GPU: k_sub_scatter with 8 keys per thread (32 registers, 2048 threads per SM): parity + bench A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "overflow_and_legacy and SUB" > gpurun_out/r2k_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_tests.log
for v in "256 16" "256 8" "512 8" "512 16"; do set -- $v
PEDK_SUB_THREADS=$1 PEDK_SUB_ITEMS=$2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2k_bench_$1_$2.log 2>&1; echo "bench $v rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2k_bench_*.log')):
    l=[x for x in open(f) if x.startswith('{')][-1]; d=json.loads(l)
    k=d['roofline']['all_kernels']
    print(f, round(d['ms_per_step'],2), 'sub_scatter', round(k['sub_scatter']['ms_per_step'],2), 'join_sub', round(k['join_sub']['ms_per_step'],2), 'count', round(k['bucket_count']['ms_per_step'],2), 'parity', d['parity']['match'])
PY
