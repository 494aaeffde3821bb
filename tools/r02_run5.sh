#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest5.log 2>&1; echo "pytest rc $?"; tail -6 gpurun_out/r02_pytest5.log
timeout -k 10 900 python bench.py > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err; echo "bench rc $?"; tail -c 800 gpurun_out/r02_bench5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench5.json'))
print(d['value'], d['e2e']['value'], d['parity_check'])
print(json.dumps(d['likelihood'])[:1500])
print({k:(round(v['value']),round(v['ms_per_step'],2)) for k,v in d['configs'].items()})
PY
