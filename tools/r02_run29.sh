#!/bin/bash
mkdir -p gpurun_out
GTTS_PROFILE_TRAIN=1 timeout -k 10 600 python tools/gpu_diag.py profile_vjp > gpurun_out/r02_profile_train.txt 2>&1; echo rc $?
head -40 gpurun_out/r02_profile_train.txt | cut -c1-120
