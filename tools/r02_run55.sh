#!/bin/bash
# session 2: incremental tile walk in the per-tap kernel's TMA producer
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_decoder.py tests/test_gpu_vocoder.py -m gpu -x -q > gpurun_out/r02_s2_pytest6.log 2>&1; echo "gpu tests rc $?"; tail -2 gpurun_out/r02_s2_pytest6.log | cut -c1-300
timeout -k 10 300 python tools/gpu_diag.py profile > gpurun_out/r02_s2_profile5.txt 2>&1; echo "profile rc $?"
grep -E "total|conv1x1|conv3x3s2" gpurun_out/r02_s2_profile5.txt | head -18
