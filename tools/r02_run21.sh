#!/bin/bash
mkdir -p gpurun_out
GTTS_ENC_PROFILE=1 timeout -k 10 600 python - > gpurun_out/r02_encoder_prof.log 2>&1 <<'PY'
import importlib, torch, os
pkg = importlib.import_module("grad-tts_b200")
te = importlib.import_module("grad-tts_b200.model.text_encoder")
cfg = pkg.synth.TEXT_ENCODER_CONFIGS["ref"]
enc = te.TextEncoder(**cfg); enc.load_state_dict(pkg.synth.make_text_encoder_state_dict(cfg, 1)); enc = enc.cuda().eval()
for B, T in [(1, 100), (128, 200)]:
    x, l, _ = pkg.synth.make_text_inputs(cfg, B, T, seed=3, ragged=False)
    for _ in range(2):
        enc(x.cuda(), l.cuda())
    torch.cuda.synchronize()
PY
echo rc $?; tail -40 gpurun_out/r02_encoder_prof.log
