#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_text_encoder.py tests/test_gpu_vocoder.py -m gpu -q > gpurun_out/r02_enc_tests.log 2>&1; echo "tests rc $?"; tail -60 gpurun_out/r02_enc_tests.log | cut -c1-400
