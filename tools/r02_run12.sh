#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest12.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02_pytest12.log
timeout -k 10 300 python bench.py --workload C1 --no-sub --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r02_bench12_c1.json 2>gpurun_out/r02_bench12_c1.err; echo "c1 rc $?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench12_c1.json')); print('C1 ms', d['ms_per_step'], 'e2e fps', d['e2e']['value'], d['roofline']['non_conv_ms'])"
timeout -k 10 300 python tools/gpu_diag.py profile > gpurun_out/r02_profile12.txt 2>&1; python tools/prof_summary.py gpurun_out/r02_profile12.txt > gpurun_out/tmp_sum.txt 2>/dev/null; head -14 gpurun_out/tmp_sum.txt
