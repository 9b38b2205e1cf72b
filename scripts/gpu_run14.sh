#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_part_scatter|k_sub_scatter|k_bucket_count' -c 3 -o gpurun_out/prof_r1e -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu14.log 2>&1; echo "ncu rc=$?"; tail -1 gpurun_out/ncu14.log | cut -c1-300
