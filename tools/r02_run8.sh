#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest8.log 2>&1; echo "pytest rc $?"; grep -E "passed|failed|assert [0-9]|FAILED" gpurun_out/r02_pytest8.log | head -20
python tools/gpu_diag.py dec_fp32 2>&1 | tail -12
