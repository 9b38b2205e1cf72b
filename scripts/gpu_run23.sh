#!/bin/bash
mkdir -p gpurun_out
for big in 2; do
PEDK_PART_BIG_TILE=$big timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench23_big$big.log 2>&1; echo "bench big=$big rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench23_big?.log')):
    l=[x for x in open(f) if x.startswith('{')][-1]
    d=json.loads(l); print(f, round(d['ms_per_step'],1), 'part_scatter', round(d['roofline']['all_kernels']['part_scatter']['ms_per_step'],2))
PY
