#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_vocoder.py -m gpu -q -x > gpurun_out/r02_voc_tests.log 2>&1; echo "tests rc $?"; tail -6 gpurun_out/r02_voc_tests.log | cut -c1-300
timeout -k 10 600 python tools/gpu_diag.py vocoder > gpurun_out/r02_vocoder_diag.log 2>&1; echo "diag rc $?"; head -9 gpurun_out/r02_vocoder_diag.log | cut -c1-160; grep "resblocks.[0-9]*.c1\|forward B=1 " gpurun_out/r02_vocoder_diag.log | awk 'NR%3==1' | head -16
