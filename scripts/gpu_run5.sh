#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=6 > gpurun_out/t_gpu5.log 2>&1; echo "pytest gpu rc=$?"
tail -8 gpurun_out/t_gpu5.log
timeout 300 python scripts/sort_bench.py 268435456 40 2>&1 | tee gpurun_out/sort_bench5.log | tail -5
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_full5.log 2>&1; echo "bench full rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_full5.log') if x.startswith('{')][-1]
d=json.loads(l); print(d['ms_per_step'], d['value']/1e9, d['stage_ms_per_step'], d['roofline']['frac'], d['e2e'])
PY
