#!/bin/bash
# new bucketed counting path: parity first, then a small and a full bench with traces
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/t11_parity.log 2>&1; echo "parity rc=$?"; tail -15 gpurun_out/t11_parity.log
PEDK_TRACE=1 timeout 600 python bench.py --scale 0.05 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench11_s005.log 2>&1; echo "small bench rc=$?"; grep -v "^\[pedk" gpurun_out/bench11_s005.log | tail -3
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench11_full.log 2>&1; echo "full bench rc=$?"; tail -2 gpurun_out/bench11_full.log
PEDK_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench11_full_trace.log 2>&1; echo "trace rc=$?"
