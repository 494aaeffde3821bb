#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_loss.py -m gpu -q > gpurun_out/r02_loss_tests.log 2>&1; echo "loss tests rc $?"; tail -8 gpurun_out/r02_loss_tests.log | cut -c1-300
