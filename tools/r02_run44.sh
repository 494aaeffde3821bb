#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_vocoder.py -m gpu -q -x > gpurun_out/r02_voc_tests.log 2>&1; echo "voc tests rc $?"; tail -3 gpurun_out/r02_voc_tests.log | cut -c1-200
for v in 1 0; do
  if [ $v = 1 ]; then export GTTS_VOC_NOPDL=1; else unset GTTS_VOC_NOPDL; fi
  timeout -k 10 300 python tools/gpu_diag.py vocoder 2>&1 | grep "forward B=" 
done
