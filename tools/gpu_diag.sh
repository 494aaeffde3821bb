#!/bin/bash
# Runs every diagnostic stage under its own timeout so that a hung kernel cannot take the whole call down.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/diag_env.txt 2>&1
for st in "$@"; do
  echo "### $st" | tee -a gpurun_out/diag.log
  timeout -k 10 120 python tools/gpu_diag.py $st >> gpurun_out/diag.log 2>&1
  echo "### $st exit $?" | tee -a gpurun_out/diag.log
done
tail -c 6000 gpurun_out/diag.log
