#!/bin/bash
# round 2, final 1-GPU pass over HEAD: full GPU tests, default bench, configs3 (capacity passes), ncu launch list
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2z_tests.log 2>&1 ) 2> gpurun_out/r2z_tests_time.txt; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2z_bench.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -1 gpurun_out/r2z_bench_ref.log | cut -c1-400
timeout 1200 python bench.py --steps 2 --warmup 1 --workload configs3 --no-cpu-baseline --no-e2e > gpurun_out/r2z_bench_c3.log 2>&1; echo "bench c3 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2z_ncu.log 2>&1; echo "ncu list rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2z_bench*.log')):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]
        d=json.loads(l)
        if d.get('impl')=='reference': continue
        print(f, round(d['ms_per_step'],1), round(d['value']/1e9,2), {k:round(v,1) for k,v in d['stage_ms_per_step'].items()}, d['config']['edges'], 'e2e', d['e2e'] and (round(d['e2e']['value']/1e9,2), round(d['e2e']['ms_per_step'],1)), 'parity', d['parity'] and d['parity']['match'], 'peakGB', round(d['peak_device_bytes']/1e9,1))
        print('   ', {k:(round(v['ms_per_step'],2),v['launches']) for k,v in d['roofline']['all_kernels'].items()})
    except Exception as e: print(f, 'ERR', e); print(open(f).read()[-1500:])
PY
