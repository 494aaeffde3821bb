#!/bin/bash
mkdir -p gpurun_out
export WG_B=16 WG_T=172
timeout -k 10 300 python tools/wgrad_check.py > gpurun_out/wg_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/wg_plain.log; exit 1; }
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel" --launch-count 12 -o gpurun_out/r02_wgrad_tc -f python tools/wgrad_check.py > gpurun_out/ncu_wgtc.log 2>&1; echo "ncu rc $?"; tail -2 gpurun_out/ncu_wgtc.log
