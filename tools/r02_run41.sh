#!/bin/bash
mkdir -p gpurun_out
GTTS_HALO1D_256=1 timeout -k 10 600 python tools/gpu_diag.py vocoder > gpurun_out/r02_vocoder_diag256.log 2>&1; echo "diag rc $?"; head -9 gpurun_out/r02_vocoder_diag256.log | tail -3 | cut -c1-160; grep "resblocks.[0-2].c[12]\|ups0" gpurun_out/r02_vocoder_diag256.log | head -12
