#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "apply_epilogue and async" > gpurun_out/r02_async_kernel2.log 2>&1; echo "async kernel tests rc $?"; tail -4 gpurun_out/r02_async_kernel2.log
timeout -k 10 300 python -m pytest tests/test_gpu_decoder.py -m gpu -x -q -k "bitwise" > gpurun_out/r02_pytest14.log 2>&1; echo "bitwise tests rc $?"; tail -3 gpurun_out/r02_pytest14.log
GTTS_FUSE_ASYNC=1 timeout -k 10 300 python tools/gpu_diag.py profile > gpurun_out/r02_profile14_async.txt 2>&1
python tools/prof_summary.py gpurun_out/r02_profile14_async.txt > gpurun_out/tmp_sum.txt 2>/dev/null; head -8 gpurun_out/tmp_sum.txt
grep -E "conv3x3ga" gpurun_out/r02_profile14_async.txt | head -28
