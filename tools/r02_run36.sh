#!/bin/bash
# N-GPU: bitwise sharding test + bench through torchrun (all sub-records, weak scaling).  usage: r02_run36.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout -k 10 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_multi$N.log 2>&1; echo "multi test rc $?"; tail -3 gpurun_out/r02_multi$N.log
timeout -k 10 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r02_bench_v4_n$N.json 2> gpurun_out/r02_bench_v4_n$N.err; echo "bench n$N rc $?"; tail -c 800 gpurun_out/r02_bench_v4_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_v4_n$N.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ranks']['ms_per_step'], d['ranks']['all_gather_ms'])
print({k:(round(v['value']),round(v['ms_per_step'],2)) for k,v in d['configs'].items()})
print(d['weak'])
PY
