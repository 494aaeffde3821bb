#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python tools/gpu_diag.py encoder > gpurun_out/r02_encoder_diag.log 2>&1; echo "diag rc $?"; tail -40 gpurun_out/r02_encoder_diag.log | cut -c1-200
timeout -k 10 900 python -m pytest tests/test_gpu_text_encoder.py tests/test_gpu_vocoder.py -m gpu -q -x > gpurun_out/r02_enc_tests.log 2>&1; echo "tests rc $?"; tail -5 gpurun_out/r02_enc_tests.log | cut -c1-400
timeout -k 10 900 python -m pytest tests/test_gpu_decoder.py -m gpu -q -x -k "fp32 or golden" > gpurun_out/r02_dec_tests.log 2>&1; echo "dec tests rc $?"; tail -5 gpurun_out/r02_dec_tests.log | cut -c1-400
