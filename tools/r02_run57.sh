#!/bin/bash
# session 2 evidence: ncu launch list of the bench command with the final kernels, and --set full of the per-sample 1x1 conv (residual
# epilogue) and the new first-conv / attn_fold kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --euler 2 --no-sub --no-cpu-baseline"
$CMD > gpurun_out/r02_s2_ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02_s2_ncu_plain.log; exit 1; }
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 400 --csv --log-file gpurun_out/r02_s2_ncu_launches.csv $CMD > gpurun_out/r02_s2_ncu_launches.log 2>&1
echo "launch list rc $?"
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|first_conv_mma|attn_fold_wide" -c 7 -o gpurun_out/r02_s2_prof_thin -f $CMD > gpurun_out/r02_s2_prof_thin.log 2>&1
echo "set full rc $?"
ls -la gpurun_out | grep r02_s2_ | tail
