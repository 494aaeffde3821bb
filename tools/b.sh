#!/bin/bash
# build the shared library; print only errors; exit non-zero on failure (so `tools/b.sh && gpurun ...` is safe)
out=$(python "$(dirname "$0")/../grad-tts_b200/build.py" -v 2>&1); rc=$?
if [ $rc -ne 0 ]; then echo "$out" | grep -E "error|Error" -A4 | head -40; echo "BUILD FAILED"; exit 1; fi
echo "build ok"
