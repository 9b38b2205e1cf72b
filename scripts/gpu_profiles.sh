#!/bin/bash
# round-1 final captures: default bench (with e2e + cpu baseline), reference arm, launch list, full ncu of the top kernels
mkdir -p gpurun_out
nproc > gpurun_out/host20.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/host20.txt
( time timeout 1200 python bench.py > gpurun_out/bench20_default.log 2>&1 ) 2> gpurun_out/bench20_time.txt; echo "bench rc=$?"; tail -3 gpurun_out/bench20_time.txt
( time timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench20_reference.log 2>&1 ) 2> gpurun_out/bench20_ref_time.txt; echo "ref rc=$?"; tail -3 gpurun_out/bench20_ref_time.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches20_full.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu20a.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_radix_pass|k_part_scatter|k_sub_scatter|k_sub_hist|k_bucket_count|k_pack_fixed|k_edges|k_line_index|k_dedup' -c 14 -o gpurun_out/prof_r1g -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu20b.log 2>&1; echo "ncu full rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench20_default.log') if x.startswith('{')][-1]
d=json.loads(l); print(round(d['ms_per_step'],1), round(d['value']/1e9,2), 'e2e', d['e2e'], 'cpu', d['cpu_baseline'])
print(d['roofline']['kernel'], d['roofline']['frac'])
PY
