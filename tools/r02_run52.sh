#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q -rP -k "first_conv_on_tensor" > gpurun_out/r02_s2_pytest3a.log 2>&1; echo "first conv tests rc $?"; grep "rel-rms" gpurun_out/r02_s2_pytest3a.log | sort | uniq | head
timeout -k 10 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_s2_pytest3.log 2>&1; echo "gpu tests rc $?"; tail -4 gpurun_out/r02_s2_pytest3.log | cut -c1-300
