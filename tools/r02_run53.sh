#!/bin/bash
# session 2: streamed weight tiles kept per stage (per-sample 1x1 conv after the attention): full GPU suite + per-launch profile
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_s2_pytest5.log 2>&1; echo "gpu tests rc $?"; tail -4 gpurun_out/r02_s2_pytest4.log | cut -c1-300
timeout -k 10 300 python tools/gpu_diag.py profile > gpurun_out/r02_s2_profile4.txt 2>&1; echo "profile rc $?"
grep -E "total|conv1x1|conv3x3s2" gpurun_out/r02_s2_profile4.txt | head -40
