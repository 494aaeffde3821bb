#!/bin/bash
# session 2: full GPU suite + per-launch profile of one Euler step (wide attn_fold, residual prefetch in the conv epilogue)
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_s2_pytest1.log 2>&1; echo "gpu tests rc $?"; tail -4 gpurun_out/r02_s2_pytest1.log | cut -c1-300
timeout -k 10 300 python tools/gpu_diag.py profile > gpurun_out/r02_s2_profile1.txt 2>&1; echo "profile rc $?"
grep -E "total|attn_fold|conv1x1|first_conv|euler" gpurun_out/r02_s2_profile1.txt | head -60
