#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_part_count|k_part_scatter|k_sub_scatter|k_bucket_count' -c 4 -o gpurun_out/prof_r1d -f python bench.py --scale 0.25 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu13.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu13.log
ls -la gpurun_out/
