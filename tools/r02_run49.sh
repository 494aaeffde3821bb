#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python tools/voc_once.py > gpurun_out/voc_once.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/voc_once.log; exit 1; }
timeout -k 10 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none -k regex:"conv_tc_kernel|sum_lrelu|conv_post|conv_ffma" --launch-skip 170 --launch-count 85 --csv --log-file gpurun_out/r02_ncu_vocoder.csv python tools/voc_once.py > gpurun_out/ncu_voc.log 2>&1; echo "ncu rc $?"; tail -2 gpurun_out/ncu_voc.log
