#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_vocoder.py -m gpu -q > gpurun_out/r02_voc_tests.log 2>&1; echo "tests rc $?"; tail -12 gpurun_out/r02_voc_tests.log | cut -c1-300
