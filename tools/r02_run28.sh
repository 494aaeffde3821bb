#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python - > gpurun_out/r02_training_rec.log 2>&1 <<'PY'
import importlib, json, sys, torch
sys.path.insert(0, '.')
import bench
pkg = importlib.import_module("grad-tts_b200")
print(json.dumps(bench.training_record(pkg, torch, None, torch.device("cuda:0")), indent=1))
PY
echo rc $?; tail -40 gpurun_out/r02_training_rec.log
