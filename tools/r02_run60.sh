#!/bin/bash
# last GPU seconds of the round: C5 shape at 10 Euler steps with the code as committed (paired stores included)
mkdir -p gpurun_out
timeout -k 5 40 python bench.py --euler 10 --steps 2 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/r02_bench_v12_euler10.json 2> gpurun_out/r02_bench_v12_euler10.err; echo "bench rc $?"
cut -c1-400 gpurun_out/r02_bench_v12_euler10.json
