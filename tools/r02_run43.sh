#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  GTTS_PDL_SMALL=$v timeout -k 10 600 python bench.py --workload C1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c1_pdl$v.json 2> gpurun_out/c1_pdl$v.err; echo "pdl_small=$v rc $?"
  python - <<PY
import json
l=json.loads(open('gpurun_out/c1_pdl$v.json').read().strip().splitlines()[-1])
print('pdl_small=$v', l['ms_per_step'], l['value'], l.get('output_finite'))
PY
done
timeout -k 10 1500 python -m pytest tests/test_gpu_decoder.py -m gpu -q -x > gpurun_out/r02_dec_tests.log 2>&1; echo "dec tests rc $?"; tail -3 gpurun_out/r02_dec_tests.log | cut -c1-200
