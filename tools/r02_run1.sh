#!/bin/bash
# round-2 GPU call 1: tests, smoke, full default bench (all sub-records)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/r02_env.txt 2>&1
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc $?" | tee -a gpurun_out/r02_pytest1.log
tail -15 gpurun_out/r02_pytest1.log
timeout -k 10 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke1.log 2>&1; echo "smoke rc $?"; tail -4 gpurun_out/r02_smoke1.log
timeout -k 10 900 python bench.py > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; echo "bench rc $?"
tail -c 1500 gpurun_out/r02_bench1.err; head -c 3000 gpurun_out/r02_bench1.json
timeout -k 10 300 python tools/gpu_diag.py profile > gpurun_out/r02_profile1.txt 2>&1; tail -45 gpurun_out/r02_profile1.txt
