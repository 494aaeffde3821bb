#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python tools/enc_once.py 1x100 > gpurun_out/enc_once.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/enc_once.log; exit 1; }
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"conv1d_splitk|rel_attention" --launch-skip 30 --launch-count 12 -o gpurun_out/r02_enc_small -f python tools/enc_once.py 1x100 > gpurun_out/ncu_enc.log 2>&1; echo "ncu rc $?"; tail -3 gpurun_out/ncu_enc.log
