#!/bin/bash
# evidence run: ncu launch list of the bench command, DRAM bytes of the conv launches, --set full of the conv / fp32-split / apply kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --euler 2 --no-sub --no-cpu-baseline"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc $?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name regex:conv_tc --launch-skip 80 --launch-count 40 --csv --log-file gpurun_out/r02_ncu_conv_dram.csv $CMD > gpurun_out/r02_ncu_conv_dram.log 2>&1
echo "conv dram rc $?"
ncu --set full --clock-control none --import-source on -k regex:conv_tc_halo2 --launch-skip 60 -c 4 -o gpurun_out/r02_prof_conv $CMD > gpurun_out/r02_prof_conv.log 2>&1
echo "set full conv rc $?"
CMD2="python bench.py --steps 1 --warmup 1 --workload C1 --precision fp32 --no-sub --no-cpu-baseline"
$CMD2 > gpurun_out/r02_ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc_halo2 --launch-skip 30 -c 3 -o gpurun_out/r02_prof_conv_fp32split $CMD2 > gpurun_out/r02_prof_conv_fp32.log 2>&1
echo "set full fp32 split rc $?"
CMD3="python bench.py --steps 1 --warmup 1 --workload C1 --no-sub --no-cpu-baseline"
$CMD3 > gpurun_out/r02_ncu_plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 200 -c 200 --csv --log-file gpurun_out/r02_ncu_launches_c1.csv $CMD3 > gpurun_out/r02_ncu_launches_c1.log 2>&1
echo "c1 launch list rc $?"
ls -la gpurun_out | grep r02_ | tail -20
