#!/bin/bash
mkdir -p gpurun_out
PEDK_TRACE=2 timeout 300 python scripts/sanitize_case.py tiny c1_40x > gpurun_out/debug_tiny.log 2>&1; echo rc=$?; grep -v "^\[pedk  " gpurun_out/debug_tiny.log | tail -30
