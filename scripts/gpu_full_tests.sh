#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/t21_gpu_all.log 2>&1 ) 2> gpurun_out/t21_time.txt; echo "pytest rc=$?"; tail -6 gpurun_out/t21_gpu_all.log; tail -3 gpurun_out/t21_time.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke21.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke21.log
