#!/bin/bash
# session 2 final: full GPU suite, smoke, bench (default flags) and the reference arm of the code as committed
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_full.log 2>&1; echo "gpu tests rc $?"; tail -2 gpurun_out/r02_gpu_tests_full.log | cut -c1-300
timeout -k 10 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/r02_smoke.log
timeout -k 10 1200 python bench.py > gpurun_out/r02_bench_v11.json 2> gpurun_out/r02_bench_v11.err; echo "bench rc $?"; tail -3 gpurun_out/r02_bench_v11.err | cut -c1-300
timeout -k 10 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r02_bench_v11_ref.json 2> gpurun_out/r02_bench_v11_ref.err; echo "ref rc $?"; cut -c1-300 gpurun_out/r02_bench_v11_ref.json
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02_bench_v11.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}, l['e2e']['value'], l['roofline']['frac'], l['roofline']['whole_step_frac'])
print('non_conv', l['roofline']['non_conv_ms'])
print('C1', l['configs']['C1']['ms_per_step'], 'training', l['training'].get('bf16'), l['training'].get('vs_gpu_eager'))
print({k:round(v['value']) for k,v in l['configs'].items()})
print('pipeline', l['pipeline'].get('ms'), l['pipeline'].get('stage_ms'))
print('clocks', l.get('clocks'))
PY
