#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests/test_gpu_decoder.py -m gpu -q -x > gpurun_out/r02_dec_tests.log 2>&1; echo "dec tests rc $?"; tail -5 gpurun_out/r02_dec_tests.log | cut -c1-300
for side in 0 1; do
  GTTS_SIDE=$side timeout -k 10 600 python bench.py --workload C1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c1_side$side.json 2> gpurun_out/c1_side$side.err; echo "side=$side rc $?"
  python - <<PY
import json
l=json.loads(open('gpurun_out/c1_side$side.json').read().strip().splitlines()[-1])
print('side=$side', l['ms_per_step'], l['value'], l.get('gpu_launches'))
PY
done
