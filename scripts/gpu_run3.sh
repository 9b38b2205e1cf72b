#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=6 > gpurun_out/t_gpu3.log 2>&1; echo "pytest gpu rc=$?"
tail -8 gpurun_out/t_gpu3.log
for m in 0 1 2; do PEDK_RANK_MODE=$m timeout 300 python scripts/sort_bench.py 268435456 40 2>&1 | tee -a gpurun_out/sort_bench3.log | tail -4; done
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_full3.log 2>&1; echo "bench full rc=$?"
tail -2 gpurun_out/bench_full3.log
