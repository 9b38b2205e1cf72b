#!/bin/bash
# compute-sanitizer memcheck + racecheck (+ synccheck) on small cases through the C ABI; logs -> gpurun_out/ (copied to profiles/)
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
N=$(nvidia-smi -L | wc -l)
for tool in memcheck racecheck synccheck; do
  timeout 1500 $CS --tool $tool --print-limit 20 python scripts/sanitize_case.py tiny c1_40x ref30k > gpurun_out/r2_sanitizer_${tool}_n1.log 2>&1
  echo "$tool n1 rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_CASE|edges" gpurun_out/r2_sanitizer_${tool}_n1.log | tail -6
done
if [ "$N" -ge 2 ]; then
for tool in memcheck racecheck; do
  timeout 1500 $CS --tool $tool --target-processes all --print-limit 20 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tests/multi_rank_check.py tiny c1_40x > gpurun_out/r2_sanitizer_${tool}_n2.log 2>&1
  echo "$tool n2 rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|MULTI_RANK" gpurun_out/r2_sanitizer_${tool}_n2.log | tail -6
done
fi
