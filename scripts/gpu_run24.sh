#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_radix_pass|k_copy_rows|k_edges' -c 7 -o gpurun_out/prof_r1g_radix -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu24.log 2>&1; echo "ncu rc=$?"
grep -o '"edges": [0-9]*' gpurun_out/ncu24.log | head -1
