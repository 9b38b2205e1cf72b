#!/bin/bash
# round 2 (1 GPU): full GPU tests (default), the same under PEDK_GUARD=1, a parity subset with the TMA bulk stores,
# default bench + e2e trace, drop-in wall
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2f_tests.log 2>&1 ) 2> gpurun_out/r2f_tests_time.txt; echo "pytest rc=$?"; tail -4 gpurun_out/r2f_tests.log
PEDK_GUARD=1 timeout 1800 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2f_tests_guard.log 2>&1; echo "guard-zone test run rc=$?"; tail -3 gpurun_out/r2f_tests_guard.log
PEDK_PART_BULK=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny or adversarial or dup or other_k or many_instances or corner or c1_40x or mixed or clip or medium" > gpurun_out/r2f_tests_bulk.log 2>&1; echo "bulk parity rc=$?"; tail -3 gpurun_out/r2f_tests_bulk.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.log 2>&1; echo "bench rc=$?"
PEDK_PART_BULK=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2f_bench_bulk1.log 2>&1; echo "bench bulk rc=$?"
PEDK_BUCKET_B_EXACT=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2f_bench_bexact.log 2>&1; echo "bench B exact rc=$?"
timeout 1200 python bench.py --steps 2 --warmup 1 --workload configs3 --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench_c3.log 2>&1; echo "bench c3 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f_bench*.log')):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]
        d=json.loads(l)
        print(f, round(d['ms_per_step'],1), round(d['value']/1e9,2), {k:round(v,1) for k,v in d['stage_ms_per_step'].items()}, d['config']['edges'], 'e2e', d['e2e'] and (round(d['e2e']['value']/1e9,2), round(d['e2e']['ms_per_step'],1)), 'parity', d['parity'] and d['parity']['match'], 'peakGB', round(d['peak_device_bytes']/1e9,1))
        print('   ', {k:(round(v['ms_per_step'],2),v['launches']) for k,v in d['roofline']['all_kernels'].items()})
    except Exception as e: print(f, 'ERR', e); print(open(f).read()[-1500:])
PY
bash scripts/gpu_dropin_wall.sh
