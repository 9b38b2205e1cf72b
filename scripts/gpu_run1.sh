#!/bin/bash
# first GPU pass: unit tests, parity tests, small and full bench
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1; (nproc; free -g; lscpu | head -20) > gpurun_out/host.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --maxfail=6 > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"
tail -15 gpurun_out/t_kernels.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu --maxfail=6 > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
tail -30 gpurun_out/t_parity.log
timeout 600 python bench.py --scale 0.05 --steps 2 --warmup 1 > gpurun_out/bench_s005.log 2>&1; echo "bench small rc=$?"
tail -5 gpurun_out/bench_s005.log
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_full.log 2>&1; echo "bench full rc=$?"
tail -5 gpurun_out/bench_full.log
