#!/bin/bash
# 8 GPUs: bench through torchrun with every sub-record (C4 = 100 n-best samples over 8 GPUs, weak scaling, per-rank times)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout -k 10 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench n8 rc $?"; tail -c 800 gpurun_out/r02_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n8.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ranks'])
print({k:(round(v['value']),round(v['ms_per_step'],2)) for k,v in d['configs'].items()})
print(d['weak'])
PY
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 2 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "bench n4 rc $?"; head -c 300 gpurun_out/r02_bench_n4.json
