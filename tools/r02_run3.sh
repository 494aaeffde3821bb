#!/bin/bash
mkdir -p gpurun_out
for v in "0 0" "0 1"; do
  set -- $v
  GTTS_FUSE_EPI=$1 GTTS_FUSE_GN=$2 timeout -k 10 300 python tools/gpu_diag.py profile > gpurun_out/r02_profile3_epi$1_gn$2.txt 2>&1
  echo "=== fuse_epi=$1 fuse_gn=$2"; python tools/prof_summary.py gpurun_out/r02_profile3_epi$1_gn$2.txt | head -8
  grep -E "conv3x3_|gn_apply" gpurun_out/r02_profile3_epi$1_gn$2.txt | head -34
done
