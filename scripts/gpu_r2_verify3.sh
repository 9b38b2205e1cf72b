#!/bin/bash
# 1 GPU: f2 with the rolling-window form of the look-up (default) and the per-window form, tests + at size + ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu -k "verification or reference_as_control_file_contract or perl_drop_in" > gpurun_out/r2y_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2y_tests.log | cut -c1-300
PEDK_TRACE=1 timeout 1000 python scripts/bench_verify.py > gpurun_out/r2y_verify_full_trace.log 2>&1; echo "bench_verify trace rc=$?"
grep "verify:" gpurun_out/r2y_verify_full_trace.log | head -8
timeout 1000 python scripts/bench_verify.py > gpurun_out/r2y_verify_full.log 2>&1; echo "bench_verify rc=$?"
tail -1 gpurun_out/r2y_verify_full.log | cut -c1-2500
PEDK_PROBE=windows timeout 1000 python scripts/bench_verify.py > gpurun_out/r2y_verify_full_windows.log 2>&1; echo "bench_verify (per-window form) rc=$?"
python - <<PY
import json
a=json.loads(open('gpurun_out/r2y_verify_full.log').read().strip().splitlines()[-1]); b=json.loads(open('gpurun_out/r2y_verify_full_windows.log').read().strip().splitlines()[-1])
print('same properties and counts:', a['properties']==b['properties'] and a['genotypes']==b['genotypes'], 'device ms', a['verify_device_ms'], b['verify_device_ms'])
PY
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_probe_strings|k_read_table' -c 4 -o /tmp/prof_r2y -f python scripts/bench_verify.py > gpurun_out/r2y_ncu.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py full /tmp/prof_r2y.ncu-rep gpurun_out/r2y_ncu_full.csv; echo "summary rc=$?"
cut -d, -f1-8,10,17-21 gpurun_out/r2y_ncu_full.csv
