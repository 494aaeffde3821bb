#!/bin/bash
for shape in "16 40 860 128" "16 20 430 256" "16 80 1720 64"; do
  for dbg in 0 128 256; do
    echo "== shape $shape dbg $dbg"; GTTS_CONV_DBG=$dbg timeout 120 python tools/apply_timing.py $shape async 2>&1 | grep "us/launch"
  done
done
