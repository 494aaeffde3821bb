#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest6.log 2>&1; echo "pytest rc $?"; tail -6 gpurun_out/r02_pytest6.log
timeout -k 10 300 python tools/gpu_diag.py profile_vjp > gpurun_out/r02_profile_vjp.txt 2>&1; head -120 gpurun_out/r02_profile_vjp.txt
