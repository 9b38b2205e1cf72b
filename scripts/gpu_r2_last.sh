#!/bin/bash
# the last 1-GPU run of round 2: full GPU tests and the default bench over HEAD
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2w_tests.log 2>&1 ) 2> gpurun_out/r2w_tests_time.txt; echo "pytest rc=$?"; tail -3 gpurun_out/r2w_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2w_bench.log 2>&1; echo "bench rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2w_smoke.log 2>&1; echo "smoke rc=$?"
python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/r2w_bench.log") if x.startswith("{")][-1])
print(round(d["ms_per_step"],2), round(d["value"]/1e9,2), "stages", round(sum(d["stage_ms_per_step"].values()),2), "e2e", (round(d["e2e"]["value"]/1e9,2), round(d["e2e"]["ms_per_step"],1)), "parity", d["parity"]["match"], "cpu", d["cpu_baseline"]["value"], d["roofline"]["frac"], d["clocks"])
print({k:round(v["ms_per_step"],2) for k,v in d["roofline"]["all_kernels"].items()})
PY
