#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=6 > gpurun_out/t_gpu4.log 2>&1; echo "pytest gpu rc=$?"
tail -8 gpurun_out/t_gpu4.log
for m in 1 2; do PEDK_RANK_MODE=$m timeout 600 python bench.py --scale 0.25 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_q_mode$m.log 2>&1; echo "mode $m rc=$?"; tail -1 gpurun_out/bench_q_mode$m.log | cut -c1-200; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_q_mode$m.log') if x.startswith('{')][-1]
d=json.loads(l); print('mode $m', d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['frac'])
PY
done
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_full4.log 2>&1; echo "bench full rc=$?"
tail -1 gpurun_out/bench_full4.log
BENCH="python bench.py --scale 0.05 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 $BENCH > gpurun_out/plain4_s005.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_radix_pass|k_count_rows|k_extract|k_edges|k_pack|k_read_scan|k_line_index|k_first_of_bucket|k_digit_hist|k_copy_rows' -s 27 -c 27 -o gpurun_out/prof_r1b $BENCH > gpurun_out/ncu_full4.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/ | tail -5
