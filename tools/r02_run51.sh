#!/bin/bash
# session 2: full GPU suite + per-launch profile of one Euler step (wide attn_fold v2, first conv on mma.sync)
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_s2_pytest2.log 2>&1; echo "gpu tests rc $?"; tail -4 gpurun_out/r02_s2_pytest2.log | cut -c1-300
timeout -k 10 300 python tools/gpu_diag.py profile > gpurun_out/r02_s2_profile2.txt 2>&1; echo "profile rc $?"
grep -E "total|attn_fold|first_conv" gpurun_out/r02_s2_profile2.txt | head -60
GTTS_FIRST_CONV_MMA=0 timeout -k 10 300 python tools/gpu_diag.py profile 2>&1 | grep -E "total|first_conv"
