#!/bin/bash
# round 2, second pass: GPU tests (f1, capacity passes, kernel variants), A/B of the CTA shapes, ncu of the new kernels
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2b_tests.log 2>&1 ) 2> gpurun_out/r2b_tests_time.txt; echo "pytest rc=$?"; tail -6 gpurun_out/r2b_tests.log
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2b_bench_$name.log 2>&1; echo "bench $name rc=$?"
}
run default A=1
run bc512 PEDK_BC_THREADS=512
run sub512 PEDK_SUB_THREADS=512
run join1024 PEDK_JOIN_THREADS=1024 PEDK_JOIN_SLOTS_LOG2=14
run passes2 PEDK_COUNT_PASSES=2
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b_bench_*.log')):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]
        d=json.loads(l); print(f, round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['stage_ms_per_step'].items()}, 'parity', d['parity'] and d['parity']['match'])
        print('   ', {k:round(v['ms_per_step'],2) for k,v in d['roofline']['all_kernels'].items()})
    except Exception as e:
        print(f,'ERR',e); print(open(f).read()[-2000:])
PY
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_rows_scatter|k_join_bucket|k_sub_scatter|k_bucket_count' -c 10 -o gpurun_out/prof_r2b -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2b_ncu.log 2>&1; echo "ncu full rc=$?"
