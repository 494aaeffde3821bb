#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "apply_epilogue" > gpurun_out/r02_async_kernel.log 2>&1; echo "apply kernel tests rc $?"; tail -8 gpurun_out/r02_async_kernel.log
grep -h "us/launch" gpurun_out/r02_async_kernel.log | head -30
timeout -k 10 600 python -m pytest tests/test_gpu_decoder.py -m gpu -x -q > gpurun_out/r02_pytest10.log 2>&1; echo "decoder tests rc $?"; tail -5 gpurun_out/r02_pytest10.log
for v in 1 0; do
  GTTS_FUSE_ASYNC=$v timeout -k 10 300 python tools/gpu_diag.py profile > gpurun_out/r02_profile10_async$v.txt 2>&1
  echo "=== fuse_async=$v"; python tools/prof_summary.py gpurun_out/r02_profile10_async$v.txt > gpurun_out/tmp_sum.txt 2>/dev/null; head -10 gpurun_out/tmp_sum.txt
done
grep -E "conv3x3|gn_apply" gpurun_out/r02_profile10_async1.txt | head -34
