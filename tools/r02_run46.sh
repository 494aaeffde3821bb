#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_text_encoder.py tests/test_gpu_vocoder.py -m gpu -q -x > gpurun_out/r02_enc_tests.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/r02_enc_tests.log | cut -c1-200
timeout -k 10 300 python tools/pipeline_breakdown.py 2>&1 | grep "^rep\|Error" | tail -4
