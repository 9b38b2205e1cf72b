#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny or adversarial or dup or other_k or overflow or many_instances or corner or c1_40x or fused_call or alphabet or mixed or clip" > gpurun_out/t_quick_parity.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/t_quick_parity.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_quick_full.log 2>&1; echo "full bench rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_quick_full.log') if x.startswith('{')][-1]
d=json.loads(l); print(round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['stage_ms_per_step'].items()})
for k,v in d['roofline']['all_kernels'].items(): print(k, round(v['ms_per_step'],2), round(v['frac'],3))
PY
