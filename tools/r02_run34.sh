#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_mas.py tests/test_gpu_align.py -m gpu -q > gpurun_out/r02_mas_tests.log 2>&1; echo "mas tests rc $?"; tail -8 gpurun_out/r02_mas_tests.log | cut -c1-300
timeout -k 10 600 python - > gpurun_out/r02_mas_rec.log 2>&1 <<'PY'
import importlib, json, os, sys, torch
sys.path.insert(0, '.')
import bench
pkg = importlib.import_module("grad-tts_b200")
dev = torch.device("cuda:0")
r = bench.mas_record(pkg, torch, dev, 6540.0)
print(json.dumps({k: r[k] for k in ("device_ms", "frac_of_hbm", "host_buffers_ms", "bit_exact_vs_cpu", "cpu_reference_ms")}))
os.environ["GTTS_MAS_BLOCK"] = "1"
r = bench.mas_record(pkg, torch, dev, 6540.0)
print("block-barrier kernel:", json.dumps({k: r[k] for k in ("device_ms", "bit_exact_vs_cpu")}))
PY
echo rc $?; tail -4 gpurun_out/r02_mas_rec.log
