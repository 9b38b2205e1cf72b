#!/bin/bash
# why do identical runs give 89 / 93 / 97 / 106 ms per step?  five runs with the sampler waiting for nvidia-smi's
# first line, one with allocation tracing
mkdir -p gpurun_out
for i in 1 2 3 4 5; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2j_bench_$i.log 2>&1; echo "bench $i rc=$?"
done
PEDK_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2j_bench_trace.log 2>&1; echo "trace rc=$?"
grep -c "cudaMalloc" gpurun_out/r2j_bench_trace.log; grep "cudaMalloc" gpurun_out/r2j_bench_trace.log | tail -5
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j_bench_*.log')):
    l=[x for x in open(f) if x.startswith('{')][-1]; d=json.loads(l)
    print(f, round(d['ms_per_step'],2), 'stages', round(sum(d['stage_ms_per_step'].values()),2), d['clocks'])
PY
ncu --query-metrics 2>/dev/null | grep -i "nvl" | head -12
