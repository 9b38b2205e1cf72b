#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_files.py -x -q -m gpu -k "reference or tiny or drop_in or resume or missing" > gpurun_out/t19.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/t19.log
