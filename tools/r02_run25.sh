#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1500 python bench.py > gpurun_out/r02_bench_v3.json 2> gpurun_out/r02_bench_v3.err; echo "bench rc $?"; tail -5 gpurun_out/r02_bench_v3.err | cut -c1-300
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02_bench_v3.json').read().strip().splitlines()[-1])
print(json.dumps({k:l[k] for k in ('value','pipeline','text_encoder') if k in l}, indent=1)[:4000])
PY
