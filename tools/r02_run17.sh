#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python tools/gpu_diag.py vocoder > gpurun_out/r02_vocoder_diag.log 2>&1; echo "diag rc $?"; tail -150 gpurun_out/r02_vocoder_diag.log | cut -c1-200
timeout -k 10 900 python -m pytest tests/test_gpu_vocoder.py -m gpu -x -q > gpurun_out/r02_vocoder_tests.log 2>&1; echo "tests rc $?"; tail -30 gpurun_out/r02_vocoder_tests.log | cut -c1-300
