#!/bin/bash
# round-2 GPU call 2: GroupNorm-apply epilogue -- kernel tests under a hang guard, then the suite, bench and profile
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "apply" > gpurun_out/r02_apply_kernel.log 2>&1; echo "apply kernel tests rc $?"
tail -25 gpurun_out/r02_apply_kernel.log
timeout -k 10 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc $?"
tail -12 gpurun_out/r02_pytest2.log
timeout -k 10 300 python tools/gpu_diag.py profile > gpurun_out/r02_profile2.txt 2>&1; python tools/prof_summary.py gpurun_out/r02_profile2.txt 2>/dev/null | head -50
timeout -k 10 600 python bench.py --no-sub --no-cpu-baseline > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err; echo "bench rc $?"
tail -c 600 gpurun_out/r02_bench2.err; head -c 2500 gpurun_out/r02_bench2.json
GTTS_APPLY=0 timeout -k 10 600 python bench.py --no-sub --no-cpu-baseline > gpurun_out/r02_bench2_noapply.json 2> gpurun_out/r02_bench2_noapply.err; echo "bench(no apply) rc $?"
head -c 400 gpurun_out/r02_bench2_noapply.json
