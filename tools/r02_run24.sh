#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc $?"; tail -8 gpurun_out/r02_smoke.log | cut -c1-200
timeout -k 10 900 python -m pytest tests/test_gpu_vocoder.py tests/test_gpu_text_encoder.py -m gpu -q -x > gpurun_out/r02_voc_tests.log 2>&1; echo "tests rc $?"; tail -5 gpurun_out/r02_voc_tests.log | cut -c1-300
timeout -k 10 1500 python bench.py > gpurun_out/r02_bench_v3.json 2> gpurun_out/r02_bench_v3.err; echo "bench rc $?"; tail -5 gpurun_out/r02_bench_v3.err | cut -c1-300
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02_bench_v3.json').read().strip().splitlines()[-1])
print(json.dumps({k:l[k] for k in ('value','pipeline','text_encoder') if k in l}, indent=1)[:4000])
PY
