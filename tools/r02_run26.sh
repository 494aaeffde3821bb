#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python tools/gpu_diag.py profile > gpurun_out/r02_profile_now.txt 2>&1; echo rc $?
grep -n "profile B=1 " -A70 gpurun_out/r02_profile_now.txt | cut -c1-110
