#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_bucket_count|k_edges' -c 2 --launch-skip 1 -o gpurun_out/prof_r1f -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu17.log 2>&1; echo "ncu rc=$?"; tail -1 gpurun_out/ncu17.log | cut -c1-200
