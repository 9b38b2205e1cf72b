#!/bin/bash
# round 2 (1 GPU): what the previous call lost to the 64 MiB limit of gpurun_out -- the k = 31 golden, the default
# bench line, sanitizer logs, launch list and a full ncu capture summarised ON the box -- plus the workloads again
mkdir -p gpurun_out
( time timeout 900 python scripts/make_golden_atsize.py --workload configs2 --scale 0.333 --k 31 --out gpurun_out > gpurun_out/r2e_golden_c2.log 2>&1 ) 2> gpurun_out/r2e_golden_c2_time.txt; echo "golden rc=$?"
cp gpurun_out/atsize_*.json tests/golden/
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench.log 2>&1; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2e_bench_reference.log 2>&1; echo "reference rc=$?"
timeout 900 python bench.py --steps 3 --warmup 2 --workload configs2 --scale 0.333 --k 31 --no-cpu-baseline > gpurun_out/r2e_bench_c2s_k31.log 2>&1; echo "bench c2 k31 s0.333 rc=$?"
timeout 900 python bench.py --steps 3 --warmup 2 --workload configs2 --k 31 --no-cpu-baseline --no-e2e > gpurun_out/r2e_bench_c2_k31.log 2>&1; echo "bench c2 k31 rc=$?"
timeout 900 python bench.py --steps 3 --warmup 2 --workload configs2 --k 20 --no-cpu-baseline --no-e2e > gpurun_out/r2e_bench_c2_k20.log 2>&1; echo "bench c2 k20 rc=$?"
timeout 900 python bench.py --steps 3 --warmup 2 --workload configs4 --no-cpu-baseline --no-e2e > gpurun_out/r2e_bench_c4.log 2>&1; echo "bench c4 rc=$?"
timeout 1200 python bench.py --steps 2 --warmup 1 --workload configs3 --no-cpu-baseline --no-e2e > gpurun_out/r2e_bench_c3.log 2>&1; echo "bench c3 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2e_bench*.log')):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]
        d=json.loads(l)
        if d.get('impl')=='reference': print(f, 'reference', round(d['value']/1e6,1), 'Mbases/s', d['cpu_baseline']['cores']); continue
        print(f, round(d['ms_per_step'],1), round(d['value']/1e9,2), {k:round(v,1) for k,v in d['stage_ms_per_step'].items()}, d['config']['edges'], 'e2e', d['e2e'] and round(d['e2e']['value']/1e9,2), 'parity', d['parity'] and d['parity']['match'], 'peakGB', round(d['peak_device_bytes']/1e9,1))
        print('   ', {k:round(v['ms_per_step'],2) for k,v in d['roofline']['all_kernels'].items()})
    except Exception as e: print(f, 'ERR', e); print(open(f).read()[-1500:])
PY
bash scripts/gpu_sanitizer.sh
for t in memcheck racecheck synccheck; do tail -15 gpurun_out/r2_sanitizer_${t}_n1.log; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2e_ncu_a.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_rows_scatter|k_rows_count|k_join_bucket|k_part_scatter|k_sub_scatter|k_sub_hist|k_bucket_count|k_pack_fixed|k_line_index|k_dedup' -c 26 -o /tmp/prof_r2e -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2e_ncu_b.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py full /tmp/prof_r2e.ncu-rep gpurun_out/r2e_ncu_full.csv; python scripts/ncu_summary.py launches gpurun_out/r2e_launches.csv gpurun_out/r2e_launch_summary.csv
ls -la /tmp/prof_r2e.ncu-rep; du -sh gpurun_out
