#!/bin/bash
# session 2: paired 64-byte stores in the conv epilogue (non-statistics variants): full GPU suite + per-launch profile
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_s2_pytest7.log 2>&1; echo "gpu tests rc $?"; tail -2 gpurun_out/r02_s2_pytest7.log | cut -c1-300
timeout -k 10 120 python tools/gpu_diag.py profile > gpurun_out/r02_s2_profile6.txt 2>&1; echo "profile rc $?"
grep -E "total|conv1x1|conv3x3s2|convT" gpurun_out/r02_s2_profile6.txt | head -20
