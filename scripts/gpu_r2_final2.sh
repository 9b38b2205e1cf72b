#!/bin/bash
# round 2, last 1-GPU pass over HEAD: smoke, full GPU tests, default bench + CPU arm, the other BASELINE workloads,
# drop-in wall clock
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2q_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2q_smoke.log
( time timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2q_tests.log 2>&1 ) 2> gpurun_out/r2q_tests_time.txt; echo "pytest rc=$?"; tail -3 gpurun_out/r2q_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2q_bench.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2q_bench_reference.log 2>&1; echo "bench ref rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 --workload configs2 --scale 0.333 --k 31 --no-cpu-baseline > gpurun_out/r2q_bench_c2s_k31.log 2>&1; echo "bench c2 k31 s0.333 rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 --workload configs2 --k 31 --no-cpu-baseline --no-e2e > gpurun_out/r2q_bench_c2_k31.log 2>&1; echo "bench c2 k31 rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 --workload configs2 --k 20 --no-cpu-baseline --no-e2e > gpurun_out/r2q_bench_c2_k20.log 2>&1; echo "bench c2 k20 rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 --workload configs4 --no-cpu-baseline --no-e2e > gpurun_out/r2q_bench_c4.log 2>&1; echo "bench c4 rc=$?"
timeout 1200 python bench.py --steps 3 --warmup 3 --workload configs3 --no-cpu-baseline --no-e2e > gpurun_out/r2q_bench_c3.log 2>&1; echo "bench c3 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2q_bench*.log')):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]
        d=json.loads(l)
        if d.get('impl')=='reference': print(f, 'reference', round(d['value']/1e6,1), 'Mbases/s', d['cpu_baseline']['cores']); continue
        print(f, round(d['ms_per_step'],1), round(d['value']/1e9,2), {k:round(v,1) for k,v in d['stage_ms_per_step'].items()}, d['config']['edges'], 'e2e', d['e2e'] and (round(d['e2e']['value']/1e9,2), round(d['e2e']['ms_per_step'],1)), 'parity', d['parity'] and d['parity']['match'], 'peakGB', round(d['peak_device_bytes']/1e9,1), d['clocks'])
    except Exception as e: print(f, 'ERR', e); print(open(f).read()[-1500:])
PY
bash scripts/gpu_dropin_wall.sh 2>&1 | tail -12
