#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python tools/mas_once.py > gpurun_out/mas_once.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/mas_once.log; exit 1; }
timeout -k 10 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_mas.csv python tools/mas_once.py > gpurun_out/ncu_mas.log 2>&1; echo "ncu rc $?"
grep -v "^==" gpurun_out/r02_ncu_mas.csv | cut -d, -f5,10-15 | tail -14
