#!/usr/bin/env python
"""Benchmark of the Grad-TTS hot path (reverse-diffusion mel decoder) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C5|C3|C1] [--euler N]

A "step" is one full pass of the hot path over one batch: `Diffusion.forward` (all n Euler steps) on the
workload's synthetic batch.  Metric (BASELINE.json): decoder mel-frames/s = B*T / time; RTF = 86.13 / (frames/s)
(reference inference.py:91).  Workload at N=1 is BASELINE config 5 (batch 128 x 1720 frames, 100 Euler steps); with
N ranks the 128 utterances are split across the ranks (no data-path collective; one all-gather of the mels).
Prints ONE JSON line on rank 0.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# The CPU legs (--impl reference, cpu_baseline) must see every host core: torchrun exports OMP_NUM_THREADS=1 for its
# workers, which caps the OpenMP/MKL pools before torch is imported.
if "--impl" in sys.argv and "reference" in sys.argv or os.environ.get("OMP_NUM_THREADS") == "1":
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

WORKLOADS = {
    # name: (n_spks, global batch, T, euler steps)
    "C5": (1, 128, 1720, 100),     # long-form throughput sweep (BASELINE.json configs[4]) -- the headline
    "C3": (247, 32, 800, 50),      # Libri-TTS multispeaker
    "C4": (1, 100, 400, 10),       # n-best: 100 samples of one utterance
    "C1": (1, 1, 400, 10),         # LJSpeech single utterance (latency)
}
FLOP_PER_FRAME_STEP = {1: 134.154e6, 247: 134.257e6}      # SURVEY 8d / BASELINE.md section 3


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1412.2), d.get("hbm_gbs", 6540.2), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_leg(n_spks, T, euler, budget_s=20.0):
    """The reference's CPU path restated (oracle/decoder_oracle.py, torch fp32, all host threads) on a bounded
    sample of the workload: B_s samples x 1 Euler step at the workload's T; frames/s at `euler` steps =
    B_s*T / (t_step * euler).  (Every Euler step costs the same and samples are independent.)"""
    import torch
    from oracle import decoder_oracle
    pkg = importlib.import_module("grad-tts_b200")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = pkg.synth.make_decoder_state_dict(n_spks, seed=0, g=0.05)
    bs = 2
    z, mask, mu, spk, _ = pkg.synth.make_inputs(bs, T, n_spks, seed=1, ragged=False)
    with torch.no_grad():
        decoder_oracle.reverse_diffusion(sd, z[:1, :, :64].contiguous(), mask[:1, :, :64].contiguous(),
                                         mu[:1, :, :64].contiguous(), 1, False, None if spk is None else spk[:1], n_spks)
        t0 = time.perf_counter()
        reps = 0
        while True:
            decoder_oracle.reverse_diffusion(sd, z, mask, mu, 1, False, spk, n_spks)
            reps += 1
            dt = time.perf_counter() - t0
            if dt > budget_s or reps >= 3:
                break
    t_step = dt / reps
    fps = bs * T / (t_step * euler)
    return {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{bs} samples x {T} frames x 1 Euler step, {reps} reps ({t_step:.2f} s/step), "
                      f"scaled to {euler} Euler steps; oracle/decoder_oracle.py (torch fp32 CPU restatement of "
                      f"model/diffusion.py) with {cores} threads"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C5", choices=sorted(WORKLOADS))
    ap.add_argument("--euler", type=int, default=0, help="override the number of Euler steps")
    ap.add_argument("--chunk", type=int, default=64)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_spks, B, T, euler = WORKLOADS[args.workload]
    if args.euler:
        euler = args.euler
    config = {"workload": f"{args.workload}: Grad-TTS decoder n_spks={n_spks}, batch {B} x {T} frames, {euler} Euler steps, "
                          f"random-init weights, synthetic mu/z, all-ones mask",
              "global_batch": B, "frames": T, "euler_steps": euler, "parallelism": f"batch-sharded x{world}",
              "l2": "inputs+workspace larger than L2 (126 MB); no explicit flush"}

    # ------------------------------------------------------------------ reference arm: CPU oracle, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return
        total = 0.0
        res = None
        for _ in range(max(1, args.warmup // 3)):
            cpu_reference_leg(n_spks, T, euler, budget_s=2.0)
        vals = []
        for _ in range(args.steps):
            res = cpu_reference_leg(n_spks, T, euler, budget_s=15.0)
            vals.append(res["value"])
        v = sum(vals) / len(vals)
        res["value"] = v
        line = {"impl": "reference", "metric": "decoder_mel_frames_per_sec", "value": v, "unit": "frames/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * B * T / v, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "rtf": 86.1328125 / v, "cpu_baseline": res,
                "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("grad-tts_b200")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a B200; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    lo, hi = pkg.dist.shard_bounds(B, world, rank)
    Bl = hi - lo
    sd = pkg.synth.make_decoder_state_dict(n_spks, seed=0, g=0.05)
    dec = pkg.Diffusion(80, 64, n_spks, 64, 0.05, 20.0, 1000)
    dec.load_state_dict(sd)
    dec = dec.to(dev)
    dec.precision = args.precision
    dec.estimator.max_chunk = args.chunk
    z, mask, mu, spk, _ = pkg.synth.make_inputs(B, T, n_spks, seed=1, ragged=False)
    zl, maskl, mul = (t[lo:hi].contiguous() for t in (z, mask, mu))
    spkl = spk[lo:hi].contiguous() if spk is not None else None
    zd, maskd, mud = zl.to(dev), maskl.to(dev), mul.to(dev)
    spkd = spkl.to(dev) if spkl is not None else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_step():
        y = dec(zd, maskd, mud, euler, False, spkd)
        if world > 1:
            y = pkg.dist.all_gather_batch(y, B)
        return y

    for _ in range(args.warmup):
        y = one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    launches = 0
    for _ in range(args.steps):
        y = one_step()
        launches += dec.estimator.launches_last_call()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    ms_per_step = ms_total / args.steps
    value = B * T / (ms_per_step * 1e-3)
    finite = bool(torch.isfinite(y).all())
    if not finite:
        raise SystemExit("bench: the sampler output contains non-finite values -- the measurement is invalid")

    # ---- e2e: host (pinned) buffers through the C-ABI host entry point, copies inside the timed region
    zp, mp, mup = zl.pin_memory(), maskl.pin_memory(), mul.pin_memory()
    spkp = spkl.pin_memory() if spkl is not None else None
    outp = torch.empty_like(zl).pin_memory()
    dec.reverse_diffusion_host(zp, mp, mup, euler, spkp, outp)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dec.reverse_diffusion_host(zp, mp, mup, euler, spkp, outp)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = B * T / (float(e2e_s.item()) / args.steps)
    h2d = (zl.numel() + mul.numel() + maskl.numel() + (spkl.numel() if spkl is not None else 0)) * 4
    d2h = zl.numel() * 4

    # ---- roofline of the dominant kernel (the tcgen05 implicit-GEMM conv), measured live with CUDA events
    roof = None
    if rank == 0:
        import ctypes
        buf = ctypes.create_string_buffer(1 << 17)
        h = dec.estimator._get_handle()
        flags = 1 if args.precision == "fp32" else 0
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            rc = pkg._lib.load().gtts_decoder_profile_step(h, Bl, T, flags, 3, buf, len(buf), ctypes.c_void_p(st))
        pkg._lib.check(rc, "profile_step")
        rep = json.loads(buf.value.decode())
        conv = [o for o in rep["ops"] if o["is_conv"]]
        t_conv = sum(o["ms"] for o in conv) * 1e-3
        t_all = sum(o["ms"] for o in rep["ops"]) * 1e-3
        f_conv = sum(o["flops"] for o in conv)
        tf_peak, hbm_peak, which = peaks()
        ach = f_conv / t_conv / 1e12
        # DRAM bytes per launch from the committed ncu capture of this command (profiles/r01_conv_dram_v11.json: 40 conv
        # launches of one Euler step at chunk 64 x 1720); only quoted when this run uses the same chunk shape
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_conv_dram_v11.json")))
            if rep["B"] == tj.get("chunk_batch", 64) and T == 1720 and tj["launches"] == len(conv):
                traffic, traffic_src = tj["traffic_bytes_per_launch"], "profiles/r01_conv_dram_v11.json (ncu dram__bytes_read+write)"
        except (OSError, KeyError, ValueError):
            pass
        roof = {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM convolutions (conv_tc / conv_tc_halo / conv_tc_halo2, all launches of one Euler step)",
                "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak, "traffic": traffic,
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": sum(o.get("bytes", 0) for o in conv) / len(conv),
                "peak_source": f"{which} (bf16_tflops_sustained)", "launches_per_step": len(conv),
                "avg_launch_ms": 1e3 * t_conv / len(conv), "conv_share_of_step": t_conv / t_all,
                "chunk_batch": rep["B"],
                "non_conv_ms": {k: sum(o["ms"] for o in rep["ops"] if o["name"].startswith(k))
                                for k in ("gn_apply", "attn_xk", "attn_ctx", "attn_merge", "attn_fold", "first_conv", "euler", "temb")}}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_reference_leg(n_spks, T, euler, budget_s=15.0)

    if rank == 0:
        fs = value * euler
        line = {"metric": "decoder_mel_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": config, "rtf": 86.1328125 / value, "frame_steps_per_sec": fs,
                "model_tflops": fs * FLOP_PER_FRAME_STEP.get(n_spks, 134.154e6) / 1e12,
                "output_finite": finite, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": "gtts_decoder_reverse_diffusion_host (C ABI, pinned host buffers)"},
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
