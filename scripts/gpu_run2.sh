#!/bin/bash
# second GPU pass: all gpu tests, host-overhead trace, launch list and full ncu captures at reduced scale
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=6 > gpurun_out/t_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -15 gpurun_out/t_gpu.log
PEDK_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/trace_full.log 2>&1; echo "trace rc=$?"
tail -3 gpurun_out/trace_full.log | cut -c1-600
BENCH="python bench.py --scale 0.05 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 $BENCH > gpurun_out/plain_s005.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_s005.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 300 $BENCH > gpurun_out/plain2_s005.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_radix_pass|k_rle|k_extract|k_edges|k_expand|k_pack|k_read_scan|k_line_index|k_first_of_bucket|k_digit_hist' -s 30 -c 30 -o gpurun_out/prof_r1a $BENCH > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/
