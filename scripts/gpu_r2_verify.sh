#!/bin/bash
# 1 GPU: full GPU tests (the result's edge records moved to pinned memory; row f2), bench with the clock
# sampler at 100 ms and at 500 ms, a stage trace of one resident step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2v_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2v_tests.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2v_bench.log 2>&1; echo "bench rc=$?"
PEDK_BENCH_SMI_MS=500 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2v_bench_smi500.log 2>&1; echo "bench rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2v_bench_again.log 2>&1; echo "bench rc=$?"
PEDK_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2v_trace.log 2>&1; echo "trace rc=$?"
tail -32 gpurun_out/r2v_trace.log | cut -c1-160
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2v_bench*.log')):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]
        d=json.loads(l)
        print(f, round(d['ms_per_step'],2), round(d['value']/1e9,2), 'stages', round(sum(d['stage_ms_per_step'].values()),2), 'e2e', d['e2e'] and (round(d['e2e']['value']/1e9,2), round(d['e2e']['ms_per_step'],1)), 'parity', d['parity'] and d['parity']['match'], d['clocks'])
    except Exception as e: print(f, 'ERR', e); print(open(f).read()[-1500:])
PY
