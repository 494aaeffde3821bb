#!/usr/bin/env python
"""Benchmark of the Grad-TTS hot path (reverse-diffusion mel decoder + MAS) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C5|C3|C4|C1] [--euler N]

A "step" is one full pass of the hot path over one batch: `Diffusion.forward` (all n Euler steps) on the
workload's synthetic batch.  Metric (BASELINE.json): decoder mel-frames/s = B*T / time; RTF = 86.13 / (frames/s)
(reference inference.py:91).  Workload at N=1 is BASELINE config 5 (batch 128 x 1720 frames, 100 Euler steps); with
N ranks the 128 utterances are split across the ranks (no data-path collective; one all-gather of the mels).
Prints ONE JSON line on rank 0.  With the default workload the line also carries, as sub-records measured in the same
run: every other BASELINE config (`configs`: C1, C3, C4, C5 at 10/50/100 Euler steps), MAS at config 2 (`mas`), the
fp32-tolerance mode (`fp32_strict`), the reference's PyTorch ops run eagerly on this GPU (`gpu_eager_baseline`, cuDNN TF32 and
bf16 autocast), a weak-scaling record at N > 1 (`weak`), a variable-length serving pattern (`varlen`), per-rank timings
(`ranks`) and a one-Euler-step error of the timed batch against the CPU oracle (`parity_check`).
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
CPU_SAMPLE_BATCH = 4                                       # reduced batch of 4-8 at the workload's T


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
        sm, smax, reasons, pw = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
                pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


def cpu_reference_leg(n_spks, T, euler, budget_s=20.0, want_output=False):
    """The reference's CPU path restated (oracle/decoder_oracle.py, torch fp32, all host threads) on a bounded
    sample of the workload: B_s samples x 1 Euler step at the workload's T; frames/s at `euler` steps =
    B_s*T / (t_step * euler).  (Every Euler step costs the same and samples are independent.)"""
    import torch
    from oracle import decoder_oracle
    pkg = importlib.import_module("grad-tts_b200")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = pkg.synth.make_decoder_state_dict(n_spks, seed=0, g=0.05)
    B = WORKLOADS_BY_SHAPE.get((n_spks, T), CPU_SAMPLE_BATCH)
    bs = min(CPU_SAMPLE_BATCH, B)
    # the first bs utterances of the workload's own batch (the generator stream depends on the batch size, so draw the full batch)
    z, mask, mu, spk, _ = pkg.synth.make_inputs(B, T, n_spks, seed=1, ragged=False)
    z, mask, mu = z[:bs].contiguous(), mask[:bs].contiguous(), mu[:bs].contiguous()
    spk = spk[:bs].contiguous() if spk is not None else None
    y = None
    with torch.no_grad():
        decoder_oracle.reverse_diffusion(sd, z[:1, :, :64].contiguous(), mask[:1, :, :64].contiguous(),
                                         mu[:1, :, :64].contiguous(), 1, False, None if spk is None else spk[:1], n_spks)
        t0 = time.perf_counter()
        reps = 0
        while True:
            y = decoder_oracle.reverse_diffusion(sd, z, mask, mu, 1, False, spk, n_spks)
            reps += 1
            dt = time.perf_counter() - t0
            if dt > budget_s or reps >= 3:
                break
    t_step = dt / reps
    fps = bs * T / (t_step * euler)
    res = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
           "sample": f"{bs} samples x {T} frames x 1 Euler step, {reps} reps ({t_step:.2f} s/step), "
                     f"scaled to {euler} Euler steps; oracle/decoder_oracle.py (torch fp32 CPU restatement of "
                     f"model/diffusion.py) with {cores} threads"}
    return (res, y) if want_output else res


WORKLOADS_BY_SHAPE = {(v[0], v[2]): v[1] for v in WORKLOADS.values()}


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
    ap.add_argument("--no-sub", action="store_true", help="main record only (no sub-records for the other configs)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_spks, B, T, euler = WORKLOADS[args.workload]
    if args.euler:
        euler = args.euler

    def config_of(name, n_spks_, B_, T_, euler_):
        return {"workload": f"{name}: Grad-TTS decoder n_spks={n_spks_}, batch {B_} x {T_} frames, {euler_} Euler steps, "
                            f"random-init weights, synthetic mu/z, all-ones mask",
                "global_batch": B_, "frames": T_, "euler_steps": euler_, "parallelism": f"batch-sharded x{world}",
                "l2": "inputs+workspace larger than L2 (126 MB); no explicit flush"}

    config = config_of(args.workload, n_spks, B, T, euler)

    # ------------------------------------------------------------------ reference arm: CPU oracle, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return
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
    import ctypes
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("grad-tts_b200")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a B200; there is no CPU fallback")
    if world > B:
        raise SystemExit(f"bench: workload {args.workload} has {B} item(s); it cannot be sharded over {world} ranks")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tf_peak, hbm_peak, which_peak = peaks()

    decoders = {}

    def get_decoder(n_spks_, precision="bf16"):
        if n_spks_ not in decoders:
            sd_ = pkg.synth.make_decoder_state_dict(n_spks_, seed=0, g=0.05)
            d_ = pkg.Diffusion(80, 64, n_spks_, 64, 0.05, 20.0, 1000)
            d_.load_state_dict(sd_)
            decoders[n_spks_] = d_.to(dev)
        d_ = decoders[n_spks_]
        d_.precision = precision
        d_.estimator.max_chunk = args.chunk
        return d_

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def gather_floats(v):
        """[v of rank 0, v of rank 1, ...] on every rank."""
        t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
        if world == 1:
            return [float(v)]
        out = torch.empty(world, device=dev, dtype=torch.float64)
        dist.all_gather_into_tensor(out, t)
        return [float(x) for x in out.cpu()]

    def shard_inputs(n_spks_, B_, T_, seed=1, tile_one_mu=False):
        z, mask, mu, spk, _ = pkg.synth.make_inputs(B_, T_, n_spks_, seed=seed, ragged=False)
        if tile_one_mu:                                   # n-best: one utterance's mu, B different z (BASELINE config 4)
            mu = mu[:1].expand(B_, -1, -1).contiguous()
            g = torch.Generator().manual_seed(seed + 1000)
            z = mu + torch.randn(mu.shape, generator=g) / 1.5
        lo, hi = pkg.dist.shard_bounds(B_, world, rank)
        host = [t[lo:hi].contiguous() for t in (z, mask, mu)] + [spk[lo:hi].contiguous() if spk is not None else None]
        devt = [t.to(dev) if t is not None else None for t in host]
        return host, devt, hi - lo

    def timed_decoder(dec, devt, B_, euler_, steps, warmup, out_buf=None):
        """Device-timed throughput of `steps` calls (max over ranks), plus this rank's split into compute and all-gather."""
        zd, maskd, mud, spkd = devt

        def one(ev=None):
            y = dec(zd, maskd, mud, euler_, False, spkd)
            if ev is not None:
                ev[0].record()
            if world > 1:
                y = pkg.dist.all_gather_batch(y, B_, out=out_buf)
            return y

        y = None
        for _ in range(warmup):
            y = one()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        mids = [[torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)] for _ in range(steps)]
        launches = 0
        barrier()
        e0.record()
        for i in range(steps):
            y = one(mids[i])
            mids[i][1].record()
            launches += dec.estimator.launches_last_call()
        e1.record()
        barrier()
        ms_local = e0.elapsed_time(e1)
        gather_ms = sum(a.elapsed_time(b) for a, b in mids) / steps
        per_rank = gather_floats(ms_local / steps)
        return y, max(per_rank), per_rank, gather_floats(gather_ms), launches

    def record(name, value, ms, n_spks_, B_, T_, euler_, extra=None):
        fs = value * euler_
        r = {"workload": name, "value": value, "unit": "frames/s", "ms_per_step": ms, "rtf": 86.1328125 / value,
             "global_batch": B_, "frames": T_, "euler_steps": euler_, "frame_steps_per_sec": fs,
             "model_tflops": fs * FLOP_PER_FRAME_STEP.get(n_spks_, 134.154e6) / 1e12}
        r["frac_of_bf16_peak_x_gpus"] = r["model_tflops"] / (tf_peak * world)
        if extra:
            r.update(extra)
        return r

    # ================================================================== main record
    dec = get_decoder(n_spks, args.precision)
    host, devt, Bl = shard_inputs(n_spks, B, T, tile_one_mu=(args.workload == "C4"))
    out_buf = torch.empty((B, 80, T), dtype=torch.float32, device=dev) if world > 1 else None
    sampler = ClockSampler(local_rank)
    sampler.start()
    y, ms_per_step, per_rank_ms, per_rank_gather, launches = timed_decoder(dec, devt, B, euler, args.steps, args.warmup, out_buf)
    clocks = sampler.stop()
    value = B * T / (ms_per_step * 1e-3)
    finite = bool(torch.isfinite(y).all())
    if not finite:
        raise SystemExit("bench: the sampler output contains non-finite values -- the measurement is invalid")
    rank_clocks = gather_floats(clocks["sm_mhz"] or 0.0)

    # ---- e2e: host (pinned) buffers through the C-ABI host entry point, copies inside the timed region
    zl, maskl, mul, spkl = host
    zp, mp, mup = zl.pin_memory(), maskl.pin_memory(), mul.pin_memory()
    spkp = spkl.pin_memory() if spkl is not None else None
    outp = torch.empty_like(zl).pin_memory()
    dec.reverse_diffusion_host(zp, mp, mup, euler, spkp, outp)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dec.reverse_diffusion_host(zp, mp, mup, euler, spkp, outp)
    barrier()
    e2e_s = max(gather_floats(time.perf_counter() - t0))
    e2e_value = B * T / (e2e_s / args.steps)
    h2d = (zl.numel() + mul.numel() + maskl.numel() + (spkl.numel() if spkl is not None else 0)) * 4
    d2h = zl.numel() * 4

    # ---- roofline of the dominant kernel (the tcgen05 implicit-GEMM conv), measured live with CUDA events
    roof, workspace = None, None
    if rank == 0:
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
        ach = f_conv / t_conv / 1e12
        # DRAM bytes per launch from the committed ncu capture of this command (40 conv launches of one Euler step at chunk
        # 64 x 1720); only quoted when this run uses the same chunk shape
        traffic, traffic_src = None, None
        for tj_name in ("r02_conv_dram.json", "r01_conv_dram_v11.json"):
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", tj_name)))
                if rep["B"] == tj.get("chunk_batch", 64) and T == 1720 and tj["launches"] == len(conv):
                    traffic = tj["traffic_bytes_per_launch"]
                    traffic_src = f"profiles/{tj_name} (ncu dram__bytes_read+write)"
                    break
            except (OSError, KeyError, ValueError):
                pass
        roof = {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM convolutions (conv_tc / conv_tc_halo2, all launches of one Euler step)",
                "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak, "traffic": traffic,
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": sum(o.get("bytes", 0) for o in conv) / len(conv),
                "peak_source": f"{which_peak} (bf16_tflops_sustained)", "launches_per_step": len(conv),
                "avg_launch_ms": 1e3 * t_conv / len(conv), "conv_share_of_step": t_conv / t_all,
                "chunk_batch": rep["B"], "euler_step_ms_eager_sum": 1e3 * t_all,
                "whole_step_frac": (value * euler * FLOP_PER_FRAME_STEP.get(n_spks, 134.154e6) / 1e12) / (tf_peak * world),
                "non_conv_ms": {k: sum(o["ms"] for o in rep["ops"] if o["name"].startswith(k))
                                for k in ("gn_apply", "attn_xk", "attn_ctx", "attn_merge", "attn_fold", "first_conv", "euler", "temb")}}
        workspace = {k: rep.get(k) for k in ("workspace_bytes", "workspace_bytes_without_reuse", "pool_bytes", "plans_cached",
                                             "plans_created")}

    # ---- CPU baseline (rank 0) and the one-Euler-step parity check of the TIMED batch against it
    cpu, parity = None, None
    y1 = dec(devt[0], devt[1], devt[2], 1, False, devt[3])[:CPU_SAMPLE_BATCH].cpu() if rank == 0 else None
    if rank == 0 and not args.no_cpu_baseline:
        cpu, yref = cpu_reference_leg(n_spks, T, euler, budget_s=15.0, want_output=True)
        nb = min(yref.shape[0], y1.shape[0])
        if args.workload != "C4":                          # C4 tiles one mu: the CPU sample uses the plain inputs
            d = (y1[:nb] - yref[:nb])
            parity = {"what": f"first {nb} utterances of the timed batch, one Euler step (n_timesteps=1) inside the full chunk, "
                              f"{args.precision} kernels vs oracle/decoder_oracle.py (fp32 CPU)",
                      "max_abs_err": float(d.abs().max()), "rel_rms": float(d.pow(2).mean().sqrt() / yref[:nb].pow(2).mean().sqrt()),
                      "ref_absmax": float(yref[:nb].abs().max()),
                      "tolerance": "bf16: rel-rms <= 2e-2; fp32: max-abs <= 1e-3 (tests/test_gpu_decoder.py)"}

    line = None
    if rank == 0:
        line = record(args.workload, value, ms_per_step, n_spks, B, T, euler)
        line.pop("workload")
        line = {"metric": "decoder_mel_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": config, "rtf": line["rtf"], "frame_steps_per_sec": line["frame_steps_per_sec"],
                "model_tflops": line["model_tflops"], "output_finite": finite, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": "gtts_decoder_reverse_diffusion_host (C ABI, pinned host buffers)"},
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "parity_check": parity,
                "workspace": workspace,
                "ranks": {"ms_per_step": per_rank_ms, "all_gather_ms": per_rank_gather, "sm_mhz_median": rank_clocks,
                          "utterances_per_rank": pkg.dist.shard_counts(B, world),
                          "note": "device time of each rank for one step (CUDA events), time spent in the single all-gather "
                                  "inside it, and each rank's median SM clock under load"}}

    # ================================================================== sub-records (default workload only)
    subs = args.workload == "C5" and not args.euler and not args.no_sub and args.precision == "bf16"
    if subs:
        configs = {}
        if rank == 0:
            configs["C5@100"] = record("C5@100", value, ms_per_step, n_spks, B, T, euler, {"e2e_value": e2e_value})
        for name, wl, eu in (("C5@50", "C5", 50), ("C5@10", "C5", 10), ("C3", "C3", 50), ("C4", "C4", 10)):
            ns_, B_, T_, _ = WORKLOADS[wl]
            if world > B_:
                continue
            d_ = get_decoder(ns_)
            if wl == "C5":
                devt_, ob_ = devt, out_buf
            else:
                _, devt_, _ = shard_inputs(ns_, B_, T_, tile_one_mu=(wl == "C4"))
                ob_ = torch.empty((B_, 80, T_), dtype=torch.float32, device=dev) if world > 1 else None
            y_, ms_, prm_, _, _ = timed_decoder(d_, devt_, B_, eu, 2, 1, ob_)
            if rank == 0:
                configs[name] = record(name, B_ * T_ / (ms_ * 1e-3), ms_, ns_, B_, T_, eu,
                                       {"output_finite": bool(torch.isfinite(y_).all()), "ms_per_rank": prm_,
                                        "utterances_per_rank": pkg.dist.shard_counts(B_, world)})
            del y_
        # ---- weak scaling: the full 128 x 1720 batch on EVERY rank (N x 128 utterances in total), 100 Euler steps
        weak = None
        if world > 1:
            zw, mw, muw, spw, _ = pkg.synth.make_inputs(B, T, n_spks, seed=1, ragged=False)
            zd, maskd, mud = zw.to(dev), mw.to(dev), muw.to(dev)
            d_ = get_decoder(n_spks)

            def weak_step():
                return d_(zd, maskd, mud, euler, False, None)
            weak_step()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            weak_step()
            e1.record()
            barrier()
            prm = gather_floats(e0.elapsed_time(e1))
            if rank == 0:
                weak = record(f"C5 weak: {B} utterances per rank, no gather", world * B * T / (max(prm) * 1e-3), max(prm), n_spks,
                              world * B, T, euler, {"scaling": "weak", "ms_per_rank": prm})
            del zd, maskd, mud
        # ---- rank-0-only records: latency config, MAS, fp32 mode, GPU eager baseline, variable lengths
        if rank == 0:
            c1 = single_gpu_config(pkg, torch, get_decoder, dev, "C1", tf_peak)
            configs["C1"] = c1
            mas = mas_record(pkg, torch, dev, hbm_peak)
            fp32 = fp32_record(pkg, torch, get_decoder, dev, devt, Bl, T)
            eager = gpu_eager_record(pkg, torch, dev, configs, line)
            varlen = varlen_record(pkg, torch, get_decoder, dev)
            line["likelihood"] = likelihood_record(pkg, torch, get_decoder, dev)
            def guarded(fn, *a):                         # the records around the headline path must not take the line down
                try:
                    return fn(*a)
                except Exception as ex:
                    torch.cuda.empty_cache()
                    return {"error": repr(ex)[:300]}
            line["vocoder"] = guarded(vocoder_record, pkg, torch, dev, line["value"])
            line["text_encoder"] = guarded(text_encoder_record, pkg, torch, dev)
            line["pipeline"] = guarded(pipeline_record, pkg, torch, dev)
            line["training"] = guarded(training_record, pkg, torch, get_decoder, dev)
            line["configs"] = configs
            line["mas"] = mas
            line["fp32_strict"] = fp32
            line["gpu_eager_baseline"] = eager
            line["weak"] = weak
            line["varlen"] = varlen
        barrier()

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------- sub-records
def _event_time_ms(torch, fn, reps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def single_gpu_config(pkg, torch, get_decoder, dev, wl, tf_peak):
    """A config that does not shard (C1: one utterance): measured on rank 0's GPU, device-timed and end to end."""
    ns_, B_, T_, eu = WORKLOADS[wl]
    d_ = get_decoder(ns_)
    z, mask, mu, spk, _ = pkg.synth.make_inputs(B_, T_, ns_, seed=1, ragged=False)
    a = [t.to(dev) if t is not None else None for t in (z, mask, mu, spk)]
    ms = _event_time_ms(torch, lambda: d_(a[0], a[1], a[2], eu, False, a[3]), 20, 5)
    zp, mp, mup = z.pin_memory(), mask.pin_memory(), mu.pin_memory()
    outp = torch.empty_like(z).pin_memory()
    d_.reverse_diffusion_host(zp, mp, mup, eu, None, outp)
    t0 = time.perf_counter()
    for _ in range(20):
        d_.reverse_diffusion_host(zp, mp, mup, eu, None, outp)
    e2e_ms = (time.perf_counter() - t0) / 20 * 1e3
    v = B_ * T_ / (ms * 1e-3)
    fs = v * eu
    return {"workload": wl, "value": v, "unit": "frames/s", "ms_per_step": ms, "rtf": 86.1328125 / v, "global_batch": B_, "frames": T_,
            "euler_steps": eu, "frame_steps_per_sec": fs, "model_tflops": fs * FLOP_PER_FRAME_STEP[1] / 1e12,
            "frac_of_bf16_peak_x_gpus": fs * FLOP_PER_FRAME_STEP[1] / 1e12 / tf_peak, "e2e_ms": e2e_ms,
            "e2e_value": B_ * T_ / (e2e_ms * 1e-3), "output_finite": bool(torch.isfinite(outp).all()),
            "note": "latency config: one utterance on one GPU (rank 0) whatever N is"}


def mas_record(pkg, torch, dev, hbm_peak):
    """BASELINE config 2: maximum_path, batch 64, text 200 x mel 1000 -- device time, host-buffer time, bit-exactness against the
    CPU restatement of the Cython kernel (oracle/mas_oracle.c) and that restatement's single-thread time."""
    from oracle import mas_oracle
    B_, tx, ty = 64, 200, 1000
    value, mask, _, _ = pkg.synth.make_mas_inputs(B_, tx, ty, seed=1234)
    v, m = value.to(dev), mask.to(dev)
    ms = _event_time_ms(torch, lambda: pkg.maximum_path(v, m, check=False), 20, 3)
    got = pkg.maximum_path(v, m).cpu()
    t0 = time.perf_counter()
    ref = mas_oracle.maximum_path(value, mask)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    lib = pkg._lib.load()
    path_h = torch.empty_like(value).pin_memory()
    vh, mh = value.pin_memory(), mask.pin_memory()
    st = torch.zeros(1, dtype=torch.int32)
    lib.gtts_mas_maximum_path_host(vh.data_ptr(), mh.data_ptr(), path_h.data_ptr(), B_, tx, ty, st.data_ptr(), dev.index or 0)
    t0 = time.perf_counter()
    for _ in range(3):
        rc = lib.gtts_mas_maximum_path_host(vh.data_ptr(), mh.data_ptr(), path_h.data_ptr(), B_, tx, ty, st.data_ptr(), dev.index or 0)
    host_ms = (time.perf_counter() - t0) / 3 * 1e3
    cells = B_ * tx * ty
    alg = 12.0 * cells                                   # value + mask read, path write (the mask multiply is fused)
    return {"workload": "C2: monotonic_align.maximum_path batch 64, text 200 x mel 1000 (ragged lengths, item 0 full size)",
            "device_ms": ms, "cells_per_sec": cells / (ms * 1e-3), "algorithmic_bytes": alg,
            "achieved_gbs": alg / (ms * 1e-3) / 1e9, "hbm_peak_gbs": hbm_peak, "frac_of_hbm": alg / (ms * 1e-3) / 1e9 / hbm_peak,
            "bound": "latency: t_y = 1000 dependent column steps per utterance (one warp carries the DP column in registers, seven warps "
                     "stage tiles), then 1000 dependent backtrack steps; one CTA per utterance (64 CTAs)",
            "host_buffers_ms": host_ms, "host_rc": int(rc), "bit_exact_vs_cpu": bool(torch.equal(got, ref)) and bool(torch.equal(path_h, ref)),
            "cpu_reference_ms": cpu_ms, "cpu_reference": "oracle/mas_oracle.c (plain-C restatement of core.pyx, 1 thread) incl. the wrapper's "
                                                         "value*mask and casts"}


def fp32_record(pkg, torch, get_decoder, dev, devt, Bl, T):
    """The fp32-tolerance mode (precision='fp32': the mode that meets max-abs 1e-3 against the fp32 reference), timed."""
    d_ = get_decoder(1, "fp32")
    ns_, B1, T1, eu1 = WORKLOADS["C1"]
    z, mask, mu, _, _ = pkg.synth.make_inputs(B1, T1, 1, seed=1, ragged=False)
    a = [t.to(dev) for t in (z, mask, mu)]
    ms1 = _event_time_ms(torch, lambda: d_(a[0], a[1], a[2], eu1), 3, 1)
    nsteps = 2
    ms5 = _event_time_ms(torch, lambda: d_(devt[0], devt[1], devt[2], nsteps, False, devt[3]), 1, 1)
    fs = Bl * T * nsteps / (ms5 * 1e-3)
    # the CUDA-core FFMA convolutions of round 1, for comparison (same tolerance, option fp32_tc=0)
    d_.estimator.set_option("fp32_tc", 0)
    ms5_ffma = _event_time_ms(torch, lambda: d_(devt[0], devt[1], devt[2], nsteps, False, devt[3]), 1, 1)
    d_.estimator.set_option("fp32_tc", 1)
    d_.precision = "bf16"
    return {"mode": d_.estimator.fp32_conv_impl(), "ffma_convs_frame_steps_per_sec": Bl * T * nsteps / (ms5_ffma * 1e-3),
            "C1": {"ms_per_step": ms1, "value": B1 * T1 / (ms1 * 1e-3), "unit": "frames/s", "euler_steps": eu1},
            "C5_shape": {"batch": Bl, "frames": T, "euler_steps_timed": nsteps, "ms": ms5, "frame_steps_per_sec": fs,
                         "frames_per_sec_at_100_steps": fs / 100.0, "model_tflops": fs * FLOP_PER_FRAME_STEP[1] / 1e12,
                         "note": "this rank's shard of the C5 batch, 2 Euler steps (every step costs the same)"}}


def gpu_eager_record(pkg, torch, dev, configs, line):
    """Second reported baseline: the reference's own PyTorch op sequence (oracle/decoder_oracle.py, a functional
    restatement of model/diffusion.py that runs on any device) executed EAGERLY on this B200 -- cuDNN/cuBLAS with torch's defaults
    (TF32 convolutions) and under bf16 autocast.  Bounded sample, scaled linearly in batch and Euler steps, like the CPU leg."""
    from oracle import decoder_oracle
    out = {"what": "torch eager on the same GPU: F.conv2d / group_norm / softplus / tanh / softmax / einsum as model/diffusion.py issues "
                   "them; no CUDA graph, default stream; torch " + torch.__version__,
           "allow_tf32": {"cudnn": bool(torch.backends.cudnn.allow_tf32), "matmul": bool(torch.backends.cuda.matmul.allow_tf32)}}
    for wl, bs, steps in (("C1", 1, 10), ("C3", 32, 2), ("C5", 16, 2)):
        ns_, B_, T_, eu = WORKLOADS[wl]
        sd = {k: v.to(dev) for k, v in pkg.synth.make_decoder_state_dict(ns_, seed=0, g=0.05).items()}
        z, mask, mu, spk, _ = pkg.synth.make_inputs(bs, T_, ns_, seed=1, ragged=False)
        a = [t.to(dev) if t is not None else None for t in (z, mask, mu, spk)]
        rec = {"sample": f"{bs} x {T_} frames x {steps} Euler steps, scaled to batch {B_} x {eu} steps"}
        for mode in ("tf32_default", "bf16_autocast"):
            def run():
                with torch.no_grad():
                    if mode == "bf16_autocast":
                        with torch.autocast("cuda", dtype=torch.bfloat16):
                            return decoder_oracle.reverse_diffusion(sd, a[0], a[1], a[2], steps, False, a[3], ns_)
                    return decoder_oracle.reverse_diffusion(sd, a[0], a[1], a[2], steps, False, a[3], ns_)
            try:
                ms = _event_time_ms(torch, run, 2, 1)
                rec[mode] = {"ms_sample": ms, "value": bs * T_ / (ms * 1e-3 / steps * eu), "unit": "frames/s"}
            except Exception as e:  # the baseline must never take the bench down
                rec[mode] = {"error": repr(e)[:200]}
                torch.cuda.empty_cache()
        best = max((rec[m].get("value", 0.0) for m in ("tf32_default", "bf16_autocast")), default=0.0)
        ours = line["value"] if wl == "C5" else (configs.get(wl) or {}).get("value")
        rec["best_eager_value"] = best
        rec["ours_value"] = ours
        rec["n_gpus_ours"] = line["n_gpus"] if wl != "C1" else 1
        rec["vs_gpu_eager"] = (ours / (best * rec["n_gpus_ours"])) if (ours and best) else None
        out[wl] = rec
        del sd, a
        torch.cuda.empty_cache()
    out["note"] = "vs_gpu_eager = our frames/s per GPU / best eager frames/s on one GPU, same config"
    return out


def likelihood_record(pkg, torch, get_decoder, dev):
    """n-best rescoring shape of the reference (n_best/config/generate_scores.yaml:3-4: 100 hypotheses per utterance, n_euler 10):
    probability-flow log-likelihood of 100 mels of 400 frames, 10 fixed steps, Hutchinson divergence.  Ours: one fused score + VJP
    call per step, state on the device (grad-tts_b200/likelihood.py).  Baseline: the reference's algorithm (oracle/likelihood_oracle.py:
    two estimator forwards + torch.autograd per step) run eagerly on this GPU on 10 hypotheses, scaled to 100."""
    from oracle import likelihood_oracle
    B_, T_, eu = 100, 400, 10
    lik = pkg.likelihood
    z, mask, mu, _, _ = pkg.synth.make_inputs(B_, T_, 1, seed=3, ragged=False)
    mu = mu[:1].expand(B_, -1, -1).contiguous()                       # one text, 100 candidate mels
    y, maskd, mud = (mu + (z - mu)).to(dev), mask.to(dev), mu.to(dev)
    g = torch.Generator().manual_seed(4)
    eps = (torch.randint(0, 2, y.shape, generator=g).float() * 2 - 1).to(dev)
    out = {"workload": f"{B_} hypotheses x {T_} frames, {eu} Euler steps of the probability-flow ODE with Hutchinson divergence "
                       "(n_best/config/generate_scores.yaml shape)"}

    class Score(torch.nn.Module):
        def __init__(self, est):
            super().__init__()
            self.estimator, self.y_mask, self.mu_y, self.spk = est, maskd, mud, None

    vals = {}
    for prec in ("fp32", "bf16"):
        d_ = get_decoder(1, prec)
        sde = lik.SPEECHSDE(0.05, 20.0, 1000, mud, None, maskd)
        fn = lik.get_likelihood_fn(sde, euler=eu)
        model = Score(d_.estimator)
        ms = _event_time_ms(torch, lambda: fn(model, y, epsilon=eps), 2, 1)
        bpd = fn(model, y, epsilon=eps)[0]
        vals[prec] = bpd
        out[prec] = {"ms": ms, "hypotheses_per_sec": B_ / (ms * 1e-3), "finite": bool(torch.isfinite(bpd).all())}
    d_.precision = "bf16"
    out["bf16_vs_fp32_rel_err_of_bpd"] = float(((vals["bf16"] - vals["fp32"]).abs() / vals["fp32"].abs()).max())
    try:
        bs = 10
        sd = {k: v.to(dev) for k, v in pkg.synth.make_decoder_state_dict(1, seed=0, g=0.05).items()}
        run = lambda: likelihood_oracle.likelihood(sd, y[:bs], maskd[:bs], mud[:bs], eu, eps[:bs])
        ms = _event_time_ms(torch, run, 1, 1)
        ref = run()[0]
        out["gpu_eager"] = {"ms_sample": ms, "sample": f"{bs} hypotheses, scaled to {B_}", "hypotheses_per_sec": bs / (ms * 1e-3),
                            "what": "reference algorithm (two estimator forwards + torch.autograd per step, cuDNN TF32) on this GPU"}
        out["vs_gpu_eager"] = {p_: out[p_]["hypotheses_per_sec"] / out["gpu_eager"]["hypotheses_per_sec"] for p_ in ("fp32", "bf16")}
        out["fp32_vs_eager_rel_err_of_bpd"] = float(((vals["fp32"][:bs] - ref).abs() / ref.abs()).max())
    except Exception as e:
        out["gpu_eager"] = {"error": repr(e)[:200]}
    return out


def pipeline_record(pkg, torch, dev):
    """Single-utterance serving latency, text to waveform (the loop body of the reference's inference.py:84-97): token ids on the
    host -> text encoder -> durations / alignment -> 10-step decoder (BASELINE C1's step count) -> HiFi-GAN -> int16 samples on the
    host.  Synthetic weights; the durations (hence the mel length) are whatever the random duration predictor says."""
    ecfg = pkg.synth.TEXT_ENCODER_CONFIGS["ref"]
    net = pkg.GradTTS(ecfg["n_vocab"], 1, 64, 192, 768, 256, 2, 6, 3, 0.1, 4, 80, 64, 0.05, 20.0, 1000)
    net.encoder.load_state_dict(pkg.synth.make_text_encoder_state_dict(ecfg, seed=1))
    net.decoder.load_state_dict(pkg.synth.make_decoder_state_dict(1, seed=0, g=0.05))
    net = net.to(dev).eval()
    vcfg = pkg.synth.VOCODER_CONFIGS["v1"]
    voc = pkg.hifigan.Generator(pkg.hifigan.AttrDict(vcfg))
    voc.load_state_dict(pkg.synth.make_vocoder_state_dict(vcfg, seed=1))
    voc = voc.to(dev).eval()
    voc.remove_weight_norm()
    x, lengths, _ = pkg.synth.make_text_inputs(ecfg, 1, 100, seed=3, ragged=False)

    def run():
        torch.manual_seed(0)
        return pkg.inference.synthesize(net, voc, x.to(dev), lengths.to(dev), n_timesteps=10)
    audio, y_dec, _ = run()
    run()
    times = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        audio, y_dec, _ = run()
        times.append((time.perf_counter() - t0) * 1e3)
    frames = int(y_dec.shape[-1])
    out = {"workload": "1 utterance, 100 tokens, 10 Euler steps, HiFi-GAN V1; host token ids -> host int16 samples (wall clock, median of 5)",
           "ms": sorted(times)[len(times) // 2], "ms_all": times, "mel_frames": frames, "audio_seconds": frames * 256 / 22050,
           "finite": bool(torch.isfinite(y_dec).all())}
    out["rtf"] = out["ms"] * 1e-3 / out["audio_seconds"]
    # where it goes (device time of each stage on its own)
    xd, ld = x.to(dev), lengths.to(dev)
    out["stage_ms"] = {"text_encoder": _event_time_ms(torch, lambda: net.encoder(xd, ld), 5, 2),
                       "vocoder": _event_time_ms(torch, lambda: voc(y_dec), 5, 2)}
    frames4 = (frames + 3) // 4 * 4                              # fix_len_compatibility (model/utils.py:13-17)
    mask = torch.ones(1, 1, frames4, device=dev)
    z = torch.randn(1, 80, frames4, device=dev)
    out["stage_ms"]["decoder_10_steps"] = _event_time_ms(torch, lambda: net.decoder(z, mask, z, 10), 5, 2)
    del net, voc
    torch.cuda.empty_cache()
    return out


def training_record(pkg, torch, get_decoder, dev):
    """Decoder part of one training step at the reference's training shape (params.py: batch_size 16, out_size 2 s -> 172 frames):
    diffusion loss forward + backward to every decoder parameter (model/diffusion.py:274-287, train.py:104-118).  Ours: the
    autograd.Function over gtts_decoder_estimator_backward.  Baseline: the reference's op sequence with torch.autograd
    (oracle/loss_oracle.py) eager on this GPU."""
    from oracle import loss_oracle
    Bt, Tt = 16, 172
    x0, mask, mu, _, _ = pkg.synth.make_inputs(Bt, Tt, 1, seed=11, ragged=True)
    a = [v.to(dev) for v in (x0, mask, mu)]
    tt = torch.rand(Bt, generator=torch.Generator().manual_seed(1)).clamp(1e-5, 1 - 1e-5).to(dev)
    z = torch.randn(Bt, 80, Tt, generator=torch.Generator().manual_seed(2)).to(dev)
    out = {"workload": f"loss_t forward + backward, {Bt} x {Tt} frames, n_spks = 1, all 172 parameter tensors", "unit": "ms per step"}
    sd = pkg.synth.make_decoder_state_dict(1, seed=0, g=0.05)
    for prec in ("bf16", "fp32"):
        dec = pkg.Diffusion(80, 64, 1, 64, 0.05, 20.0, 1000)
        dec.load_state_dict(sd)
        dec = dec.to(dev).train()
        dec.precision = prec

        def step():
            for p_ in dec.parameters():
                p_.grad = None
            loss, _ = dec.loss_t(a[0], a[1], a[2], tt, noise=z)
            loss.backward()
            return loss
        loss = step()
        ms = _event_time_ms(torch, step, 5, 2)
        out[prec] = {"ms": ms, "loss": float(loss), "grads_finite": bool(all(torch.isfinite(p_.grad).all() for p_ in dec.parameters()))}
        del dec
    sd_d = {k: v.to(dev) for k, v in sd.items()}
    eager = {}
    for mode in ("tf32_default", "bf16_autocast"):
        def estep():
            if mode == "bf16_autocast":
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return loss_oracle.loss_t_grads(sd_d, a[0], a[1], a[2], tt, z, None, 1)
            return loss_oracle.loss_t_grads(sd_d, a[0], a[1], a[2], tt, z, None, 1)
        try:
            l_e = estep()[0]
            eager[mode] = {"ms": _event_time_ms(torch, estep, 5, 2), "loss": float(l_e)}
        except Exception as ex:
            eager[mode] = {"error": repr(ex)[:200]}
            torch.cuda.empty_cache()
    out["gpu_eager_baseline"] = eager
    best = min((v["ms"] for v in eager.values() if "ms" in v), default=None)
    out["vs_gpu_eager"] = (best / out["bf16"]["ms"]) if best else None
    torch.cuda.empty_cache()
    return out


def text_encoder_record(pkg, torch, dev):
    """The step before the decoder (model/tts.py:84): TextEncoder.forward on token batches.  Ours: csrc/text_encoder.cu (fp32).
    Baseline: the reference's op sequence (oracle/text_encoder_oracle.py) run eagerly on this GPU with torch's defaults."""
    from oracle import text_encoder_oracle
    te = importlib.import_module("grad-tts_b200.model.text_encoder")
    cfg = pkg.synth.TEXT_ENCODER_CONFIGS["ref"]
    sd = pkg.synth.make_text_encoder_state_dict(cfg, seed=1)
    enc = te.TextEncoder(**cfg)
    enc.load_state_dict(sd)
    enc = enc.to(dev).eval()
    sd_d = {k: v.to(dev) for k, v in sd.items()}
    out = {"unit": "tokens/s", "note": "fp32 on the CUDA cores (the output decides integer durations); eager = F.conv1d / matmul / softmax "
                                       "as model/text_encoder.py issues them, with torch's TF32 defaults and with TF32 off (same "
                                       "arithmetic as ours)"}
    for name, Bt, Tt in (("batch128_x_200_tokens", 128, 200), ("single_utterance_100_tokens", 1, 100)):
        x, lengths, _ = pkg.synth.make_text_inputs(cfg, Bt, Tt, seed=3, ragged=False)
        xd, ld = x.to(dev), lengths.to(dev)
        mu, logw, _ = enc(xd, ld)
        ms = _event_time_ms(torch, lambda: enc(xd, ld), 5, 2)

        def eager():
            with torch.no_grad():
                return text_encoder_oracle.text_encoder_forward(sd_d, cfg, xd, ld)
        mu_e, logw_e, _ = eager()
        ms_e = _event_time_ms(torch, eager, 5, 2)
        tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        try:
            mu_f, _, _ = eager()
            ms_f = _event_time_ms(torch, eager, 5, 2)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf
        out[name] = {"ms": ms, "value": Bt * Tt / (ms * 1e-3), "launches": enc.launches_last_call(), "eager_tf32_ms": ms_e,
                     "eager_fp32_ms": ms_f, "vs_gpu_eager_tf32": ms_e / ms, "vs_gpu_eager_fp32": ms_f / ms,
                     "max_abs_mu_vs_eager_fp32": float((mu - mu_f).abs().max()), "max_abs_mu_vs_eager_tf32": float((mu - mu_e).abs().max())}
    del enc, sd_d
    torch.cuda.empty_cache()
    return out


def vocoder_record(pkg, torch, dev, decoder_fps):
    """The step after the decoder in inference.py:97: HiFi-GAN V1 generator (hifi-gan/models.py:77-118) on mels of the headline
    shape.  Ours: csrc/vocoder.cu (tcgen05 1-D convs, bf16 activations; and the strict fp32 mode).  Baseline: the reference's op
    sequence (oracle/vocoder_oracle.py = F.conv1d / conv_transpose1d / leaky_relu as models.py issues them) run eagerly on this GPU."""
    from oracle import vocoder_oracle
    cfg = pkg.synth.VOCODER_CONFIGS["v1"]
    sd = pkg.synth.make_vocoder_state_dict(cfg, seed=1)
    gen = pkg.hifigan.Generator(pkg.hifigan.AttrDict(cfg))
    gen.load_state_dict(sd)
    gen = gen.to(dev).eval()
    gen.remove_weight_norm()
    Bv, T_ = 32, 1720
    out = {"workload": f"HiFi-GAN V1 generator, {Bv} mels x {T_} frames -> {T_ * 256} samples each (22.05 kHz)", "unit": "frames/s"}
    mel = pkg.synth.make_mel(Bv, T_, seed=2)
    mel_d = mel.to(dev)
    y = gen(mel_d)
    ms = _event_time_ms(torch, lambda: gen(mel_d), 3, 1)
    # algorithmic FLOPs per mel frame of the unpadded network (conv_pre, 4 transposed convs, 12 resblocks x 6 convs, conv_post)
    c0, flops, L, ch = cfg["upsample_initial_channel"], 2.0 * 80 * cfg["upsample_initial_channel"] * 7, 1, cfg["upsample_initial_channel"]
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        flops += 2.0 * L * ch * (ch // 2) * k                    # every input position meets every tap of the transposed conv
        L, ch = L * u, ch // 2
        flops += sum(2.0 * L * ch * ch * rk * 2 * len(rd) for rk, rd in zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"]))
    flops += 2.0 * L * ch * 7
    out["gflop_per_frame"] = flops / 1e9
    # algorithmic HBM bytes per mel frame of the layer-by-layer schedule (bf16 activations; weights are negligible): per stage with
    # U = positions x channels x 2 bytes, the transposed conv reads its input and writes x and lrelu(x); a (c1, c2) pair moves
    # 2U + 4U (3U when it is the resblock's last and lrelu(x) is not needed); the 3-way sum reads 3U and writes U
    Lp, chp, abytes = 1, cfg["upsample_initial_channel"], 2.0 * (80 * 4 + cfg["upsample_initial_channel"] * 2)
    for u in cfg["upsample_rates"]:
        prev = Lp * chp * 2.0
        Lp, chp = Lp * u, chp // 2
        U = Lp * chp * 2.0
        nd = len(cfg["resblock_dilation_sizes"][0])
        abytes += prev + 2 * U + len(cfg["resblock_kernel_sizes"]) * (nd * 6 * U - U) + (len(cfg["resblock_kernel_sizes"]) + 1) * U
    abytes += Lp * chp * 2.0 + Lp * 4.0
    out["algorithmic_hbm_bytes_per_frame"] = abytes
    out["bf16"] = {"ms": ms, "value": Bv * T_ / (ms * 1e-3), "launches": gen.launches_last_call(), "output_finite": bool(torch.isfinite(y).all()),
                   "audio_seconds_per_second": Bv * T_ * 256 / 22050 / (ms * 1e-3),
                   "model_tflops": flops * Bv * T_ / (ms * 1e-3) / 1e12,
                   "frac_of_bf16_peak": flops * Bv * T_ / (ms * 1e-3) / 1e12 / peaks()[0],
                   "roofline": {"bound": "hbm", "achieved": abytes * Bv * T_ / (ms * 1e-3) / 1e9, "peak": peaks()[1], "unit": "GB/s",
                                "frac": abytes * Bv * T_ / (ms * 1e-3) / 1e9 / peaks()[1],
                                "traffic_note": "profiles/r02_ncu_vocoder.csv: 81.3 GB of DRAM traffic per 16 x 1720 frames "
                                                "(dram__bytes_read + write over the 82 conv / point-wise launches) vs the "
                                                "algorithmic figure above"}}
    # end to end with host buffers through the C ABI (H2D of the mels, D2H of the waveforms inside the timed region)
    mel_h = mel.pin_memory()
    wav_h = torch.empty(Bv, 1, T_ * 256).pin_memory()
    lib = pkg._lib.load()
    t0 = time.perf_counter()
    pkg._lib.check(lib.gtts_vocoder_forward_host(gen._handle, mel_h.data_ptr(), wav_h.data_ptr(), Bv, T_, 0), "vocoder_forward_host")
    dt = time.perf_counter() - t0
    out["bf16"]["e2e_value"] = Bv * T_ / dt
    out["bf16"]["h2d_bytes"], out["bf16"]["d2h_bytes"] = mel_h.numel() * 4, wav_h.numel() * 4
    # parity of the timed batch: first 2 utterances against the fp32 eager run below
    sd_d = {k: v.to(dev) for k, v in sd.items()}
    with torch.no_grad():
        ref = vocoder_oracle.generator_forward(sd_d, cfg, mel_d[:2])
    out["bf16"]["rel_rms_vs_eager_tf32_default"] = float(((y[:2] - ref).pow(2).mean() / ref.pow(2).mean()).sqrt())
    del y
    gen.precision = "fp32"
    gen.max_chunk = 4
    y32 = gen(mel_d[:4])
    ms32 = _event_time_ms(torch, lambda: gen(mel_d[:4]), 2, 1)
    out["fp32_strict"] = {"ms": ms32, "value": 4 * T_ / (ms32 * 1e-3), "sample": f"4 x {T_} frames",
                          "max_abs_vs_eager_tf32_default": float((y32[:2] - ref).abs().max())}
    del y32
    be = 4
    eager = {"sample": f"{be} x {T_} frames, scaled to {Bv}"}
    for mode in ("tf32_default", "bf16_autocast"):
        def run():
            with torch.no_grad():
                if mode == "bf16_autocast":
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        return vocoder_oracle.generator_forward(sd_d, cfg, mel_d[:be])
                return vocoder_oracle.generator_forward(sd_d, cfg, mel_d[:be])
        try:
            ms_e = _event_time_ms(torch, run, 2, 1)
            eager[mode] = {"ms_sample": ms_e, "value": be * T_ / (ms_e * 1e-3)}
        except Exception as e:
            eager[mode] = {"error": repr(e)[:200]}
            torch.cuda.empty_cache()
    best = max((eager[m].get("value", 0.0) for m in ("tf32_default", "bf16_autocast")), default=0.0)
    eager["best_eager_value"] = best
    out["gpu_eager_baseline"] = eager
    out["vs_gpu_eager"] = out["bf16"]["value"] / best if best else None
    # what the vocoder adds to the headline pipeline (decoder at 100 Euler steps + vocoder)
    out["decoder_plus_vocoder_frames_per_s"] = 1.0 / (1.0 / decoder_fps + 1.0 / out["bf16"]["value"])
    del gen, sd_d, mel_d
    torch.cuda.empty_cache()
    return out


def varlen_record(pkg, torch, get_decoder, dev):
    """Serving pattern: every call has a different mel length (plans are keyed on (B, T); T cannot be bucketed because GroupNorm
    and the attention softmax include the padding).  First pass builds the plans (descriptors + graph; the activation memory comes
    from the shared pool), second pass hits the plan cache."""
    d_ = get_decoder(1)
    g = torch.Generator().manual_seed(5)
    Ts = sorted({int(t) // 4 * 4 for t in torch.randint(400, 1721, (12,), generator=g)})
    Bv, eu = 8, 10
    ins = []
    for T_ in Ts:
        z, mask, mu, _, _ = pkg.synth.make_inputs(Bv, T_, 1, seed=T_, ragged=True)
        ins.append([t.to(dev) for t in (z, mask, mu)])
    before = d_.estimator.cache_info()
    res = {}
    for name in ("first_pass_builds_plans", "second_pass_cached"):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for a in ins:
            d_(a[0], a[1], a[2], eu)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res[name] = {"seconds": dt, "value": Bv * sum(Ts) / dt, "unit": "frames/s"}
    after = d_.estimator.cache_info()
    res.update({"lengths": Ts, "batch": Bv, "euler_steps": eu, "plans_created": after["plans_created"] - before["plans_created"],
                "pool_bytes_before": before["pool_bytes"], "pool_bytes_after": after["pool_bytes"]})
    return res


if __name__ == "__main__":
    main()
