#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 240 python tools/wgrad_check.py > gpurun_out/r02_wgrad_check.log 2>&1; echo "check rc $?"; tail -50 gpurun_out/r02_wgrad_check.log | cut -c1-160
