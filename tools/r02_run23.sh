#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_text_encoder.py -m gpu -q -x > gpurun_out/r02_enc_tests.log 2>&1; echo "tests rc $?"; tail -15 gpurun_out/r02_enc_tests.log | cut -c1-300
GTTS_ENC_PROFILE=1 timeout -k 10 300 python tools/enc_once.py > gpurun_out/r02_encoder_prof.log 2>&1; echo rc $?; grep -A13 "profile" gpurun_out/r02_encoder_prof.log | tail -30
timeout -k 10 600 python tools/gpu_diag.py encoder 2>&1 | grep -v "^small_m=     0\|small_m=100000" | tail -12
