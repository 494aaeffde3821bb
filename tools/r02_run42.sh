#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_vocoder.py -m gpu -q -x > gpurun_out/r02_voc_tests.log 2>&1; echo "voc tests rc $?"; tail -6 gpurun_out/r02_voc_tests.log | cut -c1-300
timeout -k 10 600 python tools/gpu_diag.py vocoder > gpurun_out/r02_vocoder_diag.log 2>&1; echo "diag rc $?"; head -9 gpurun_out/r02_vocoder_diag.log | cut -c1-160; grep "ups3\|resblocks.\(9\|10\|11\).c[12]\|conv_post\|forward B=1 " gpurun_out/r02_vocoder_diag.log | head -24
timeout -k 10 1200 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/r02_dec_tests.log 2>&1; echo "dec tests rc $?"; tail -3 gpurun_out/r02_dec_tests.log | cut -c1-200
