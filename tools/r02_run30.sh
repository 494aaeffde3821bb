#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_vjp.py -m gpu -q -x > gpurun_out/r02_train.log 2>&1; echo "train tests rc $?"; tail -12 gpurun_out/r02_train.log | cut -c1-300
GTTS_PROFILE_TRAIN=1 timeout -k 10 600 python tools/gpu_diag.py profile_vjp > gpurun_out/r02_profile_train.txt 2>&1; echo rc $?
head -12 gpurun_out/r02_profile_train.txt | cut -c1-120; grep "bwd_wgrad" gpurun_out/r02_profile_train.txt | sort -k2 -n -r | head -12
timeout -k 10 900 python - > gpurun_out/r02_training_rec.log 2>&1 <<'PY'
import importlib, json, sys, torch
sys.path.insert(0, '.')
import bench
pkg = importlib.import_module("grad-tts_b200")
print(json.dumps(bench.training_record(pkg, torch, None, torch.device("cuda:0")), indent=1))
PY
echo rc $?; grep -A3 '"bf16"\|"fp32"\|vs_gpu' gpurun_out/r02_training_rec.log | head -30
