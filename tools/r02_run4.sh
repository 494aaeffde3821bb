#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_vjp.py -m gpu -x -q > gpurun_out/r02_vjp.log 2>&1; echo "vjp tests rc $?"
tail -40 gpurun_out/r02_vjp.log
