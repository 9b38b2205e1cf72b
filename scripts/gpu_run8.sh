#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --maxfail=6 > gpurun_out/t_gpu8.log 2>&1; echo "pytest gpu rc=$?"
tail -12 gpurun_out/t_gpu8.log
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_full8.log 2>&1; echo "bench full rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_full8.log') if x.startswith('{')][-1]
d=json.loads(l); print(d['ms_per_step'], d['value']/1e9, d['stage_ms_per_step'], d['roofline']['frac'], d['e2e'], d['cpu_baseline'])
PY
BENCH="python bench.py --scale 0.05 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 $BENCH > gpurun_out/plain8_s005.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches8_s005.csv $BENCH > gpurun_out/ncu_launches8.log 2>&1
echo "launch list rc=$?"
timeout 300 $BENCH > gpurun_out/plain8b_s005.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_radix_pass|k_count_rows|k_extract|k_edges|k_pack_fixed|k_line_index|k_newline_count|k_dedup' -s 22 -c 22 -o gpurun_out/prof_r1c $BENCH > gpurun_out/ncu_full8.log 2>&1
echo "ncu full rc=$?"
