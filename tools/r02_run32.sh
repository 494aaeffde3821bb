#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests_full.log 2>&1; echo "gpu tests rc $?"; tail -12 gpurun_out/r02_gpu_tests_full.log | cut -c1-300
