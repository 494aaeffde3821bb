#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q > gpurun_out/r02_train.log 2>&1; echo "train tests rc $?"; tail -40 gpurun_out/r02_train.log | cut -c1-400
