#!/bin/bash
# round 2, first pass: GPU tests on the hash join, the at-size golden of configs[1] (oracle on the box's host),
# default bench with parity, launch list + full ncu capture of the top kernels
mkdir -p gpurun_out
nproc > gpurun_out/r2a_host.txt; free -g >> gpurun_out/r2a_host.txt
( time timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2a_tests.log 2>&1 ) 2> gpurun_out/r2a_tests_time.txt; echo "pytest rc=$?"; tail -6 gpurun_out/r2a_tests.log
( time timeout 900 python scripts/make_golden_atsize.py --workload configs1 --out tests/golden > gpurun_out/r2a_golden.log 2>&1 ) 2> gpurun_out/r2a_golden_time.txt; echo "golden rc=$?"; tail -2 gpurun_out/r2a_golden.log | cut -c1-600
cp tests/golden/atsize_*.json gpurun_out/ 2>/dev/null
( time timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.log 2>&1 ) 2> gpurun_out/r2a_bench_time.txt; echo "bench rc=$?"
python - <<PY
import json
try:
    l=[x for x in open('gpurun_out/r2a_bench.log') if x.startswith('{')][-1]
    d=json.loads(l); print(round(d['ms_per_step'],1), round(d['value']/1e9,2), {k:round(v,1) for k,v in d['stage_ms_per_step'].items()})
    for k,v in d['roofline']['all_kernels'].items(): print('   ',k, round(v['ms_per_step'],2), round(v['frac'],3))
    print('e2e', d['e2e']); print('parity', d['parity']); print('cpu', d['cpu_baseline'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r2a_bench.log').read()[-3000:])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2a_ncu_a.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_rows_scatter|k_rows_count|k_join_bucket|k_part_scatter|k_sub_scatter|k_sub_hist|k_bucket_count|k_pack_fixed|k_line_index|k_dedup' -c 16 -o gpurun_out/prof_r2a -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2a_ncu_b.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/prof_r2a.ncu-rep
