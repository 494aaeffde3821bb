#!/bin/bash
# 2-GPU: bitwise sharding test + bench through torchrun (all sub-records, weak scaling)
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout -k 10 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_multi2.log 2>&1; echo "multi test rc $?"; tail -5 gpurun_out/r02_multi2.log
timeout -k 10 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc $?"; tail -c 1500 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ranks'])
print({k:(round(v['value']),round(v['ms_per_step'],2)) for k,v in d['configs'].items()})
print(d['weak'])
PY
