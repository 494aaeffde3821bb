#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fp32_on_tensor" > gpurun_out/r02_split_kernel.log 2>&1; echo "split kernel tests rc $?"; tail -15 gpurun_out/r02_split_kernel.log
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest7.log 2>&1; echo "pytest rc $?"; tail -8 gpurun_out/r02_pytest7.log
timeout -k 10 900 python bench.py --no-cpu-baseline > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err; echo "bench rc $?"; tail -c 600 gpurun_out/r02_bench7.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench7.json'))
print(d['value'], json.dumps(d['fp32_strict']))
print(json.dumps(d['likelihood'])[:1200])
PY
