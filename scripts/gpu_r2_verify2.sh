#!/bin/bash
# 1 GPU: rows f1 / f2 tests after the host-side rewrite, f1 + f2 at configs[1] size (with and without a stage
# trace), then ncu --set full of the window look-up kernel (summarised on the box)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu -k "verification or reference or perl_drop_in or cnv" > gpurun_out/r2x_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2x_tests.log | cut -c1-300
timeout 1000 python scripts/bench_verify.py > gpurun_out/r2x_verify_full.log 2>&1; echo "bench_verify rc=$?"
tail -1 gpurun_out/r2x_verify_full.log | cut -c1-2500
PEDK_TRACE=1 timeout 1000 python scripts/bench_verify.py > gpurun_out/r2x_verify_full_trace.log 2>&1; echo "bench_verify trace rc=$?"
grep "verify:" gpurun_out/r2x_verify_full_trace.log | head -8
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_probe_windows|k_lookup_sorted' -c 3 -o /tmp/prof_r2x -f python scripts/bench_verify.py > gpurun_out/r2x_ncu.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py full /tmp/prof_r2x.ncu-rep gpurun_out/r2x_ncu_full.csv; echo "summary rc=$?"
cat gpurun_out/r2x_ncu_full.csv | cut -c1-1500
