#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_full.log 2>&1; echo "gpu tests rc $?"; tail -6 gpurun_out/r02_gpu_tests_full.log | cut -c1-300
timeout -k 10 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/r02_smoke.log
timeout -k 10 1800 python bench.py > gpurun_out/r02_bench_v4.json 2> gpurun_out/r02_bench_v4.err; echo "bench rc $?"; tail -3 gpurun_out/r02_bench_v4.err | cut -c1-300
timeout -k 10 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc $?"; cut -c1-600 gpurun_out/r02_bench_ref.json
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02_bench_v4.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}, l['e2e']['value'], l['roofline']['frac'], l['roofline']['whole_step_frac'])
print('C1', l['configs']['C1']['ms_per_step'], 'training', l['training'].get('bf16'), l['training'].get('vs_gpu_eager'))
print('pipeline', l['pipeline'].get('ms'), l['pipeline'].get('stage_ms'))
print('clocks', l.get('clocks'))
PY
