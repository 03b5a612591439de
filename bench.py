#!/usr/bin/env python
"""Benchmark of the reverse-diffusion hot path: mel frames/s of 50-step CFG reverse diffusion.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one full pass of the hot path over one batch: a 50-step text+speaker classifier-free-guided reverse
diffusion (150 U-Net evaluations per utterance) of 16 utterances x 512 frames per GPU (BASELINE.json configs[1]).
With N > 1 every rank samples its own 16 utterances (weak scaling, sharded by utterance, no collective in the step
loop); the final mels are all-gathered with NCCL inside the timed region.

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the same metric through the
public API with pinned HOST tensors (H2D of z/cond/mask/spk/noise and D2H of the mels inside the timed region);
`roofline` = the tcgen05 implicit-GEMM conv kernel (algorithmic conv FLOPs / its summed CUDA-event time, measured in a
separate profiled pass of the same workload, against the measured bf16 GEMM peak); `cpu_baseline` = the CPU port of the
reference algorithm (oracle/) on a bounded sample of the same workload.

`--impl reference` times the CPU port of the reference (oracle/unitspeech_oracle.py, pinned to the reference's outputs
by tests/golden) on the host cores: the upstream package is a research script tree whose decoder imports need stubs
 and its path is pure PyTorch, so the oracle port is the reference arm that can travel to the GPU box.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU reference arm (rank 0 only, the other ranks exit at once) is
# specified to use all the host threads it can, and OpenMP reads the variable when torch is first imported
if any(a in ("reference", "--impl=reference") for a in sys.argv) and os.environ.get("RANK", "0") == "0":
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    os.environ.pop("MKL_NUM_THREADS", None)

import torch  # noqa: E402

METRIC = "mel frames/s of 50-step CFG reverse diffusion"
UNIT = "frames/s"
N_STEPS = 50                      # diffusion steps per pass
TG, SG = 1.0, 1.0                 # text / speaker guidance scales
BATCH, FRAMES = 16, 512           # per-GPU batch (BASELINE.json configs[1])
N_FEATS, SPK = 80, 256
CONV_FLOP_PER_FRAME_EVAL = 647.27e6   # conv FLOPs per mel frame per estimator evaluation
SCALE = 1.0 / 512                 # harness scale keeping the untrained sampler O(1)


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tensor": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "tensor_burst": float(p["bf16_tflops"]),
                "hbm": float(p["hbm_gbs"]), "src": "measured"}
    return {"tensor": 1400.0, "tensor_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every 100 ms while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        busy = [s for s in self.samples if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(busy)}


def make_inputs(seed: int, B: int, T: int, n: int):
    from unitspeech_b200.synthetic import synthetic_inputs
    return synthetic_inputs(B, T, n, seed=seed, scale=SCALE)


def harness_weights(dec=None):
    """Seeded random-init weights of the reference architecture (same tensors for both arms)."""
    from unitspeech_b200 import UnitSpeech
    from unitspeech_b200.synthetic import random_init_state_dict
    if dec is None:
        dec = UnitSpeech(N_FEATS, 128, (1, 2, 4, 8), beta_min=0.05, beta_max=20, pe_scale=1000, spk_emb_dim=SPK)
    return random_init_state_dict(dec, seed=1234, out_scale=SCALE)


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference algorithm)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_sample(params, threads: int, T: int = FRAMES, diffusion_steps: int = 1):
    """One utterance x T frames x `diffusion_steps` CFG steps on the host; returns (seconds, frames/s of the 50-step job)."""
    from oracle import unitspeech_oracle as O
    torch.set_num_threads(threads)
    n = max(2, diffusion_steps)
    z, mask, cond, spk, noise = make_inputs(7, 1, T, N_STEPS)
    # run the first `n` steps of the real 50-step schedule: same per-step cost as any other step
    tb = O.schedule_tables(N_STEPS, 0.05, 20.0)
    times = O.step_times(N_STEPS)
    xt = z * mask
    tu = params["text_uncon"].repeat(1, 1, T)
    su = (params["spk_uncon"] / params["spk_uncon"].norm()).repeat(1, 1, 1)
    est = lambda x_, m_, mu_, t_, s_: O.estimator_forward(params, x_, m_, mu_, t_, s_, 128, (1, 2, 4, 8))  # noqa: E731
    t0 = time.perf_counter()
    with torch.no_grad():
        for i in range(n):
            t = times[i] * torch.ones(1)
            score = O.cfg_score(params, xt, mask, cond, t, spk, tu, su, TG, SG, est)
            xt = O.sampler_step(tb, N_STEPS - 1 - i, xt, score, noise[i], mask)
    dt = time.perf_counter() - t0
    per_step = dt / n
    return dt, T / (N_STEPS * per_step), n


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    params = harness_weights()
    vals = []
    ref_steps = 4   # CFG diffusion steps per bench step: ~4 s of CPU work each
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_sample(params, threads, diffusion_steps=2)
    t_total = 0.0
    for _ in range(args.steps):
        dt, fps, n = cpu_sample(params, threads, diffusion_steps=ref_steps)
        vals.append(fps)
        t_total += dt
    v = statistics.median(vals)
    sample = (f"1 utterance x {FRAMES} frames x {ref_steps} of the 50 CFG diffusion steps ({3 * ref_steps} U-Net evaluations) "
              f"per bench step, scaled to the 50-step job")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * t_total / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"UnitSpeech decoder, 50-step text+speaker CFG (1.0/1.0) reverse diffusion, {BATCH} utt x {FRAMES} frames per GPU",
                   "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch.distributed as dist
    from unitspeech_b200 import UnitSpeech

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this framework has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    dec = UnitSpeech(N_FEATS, 128, (1, 2, 4, 8), beta_min=0.05, beta_max=20, pe_scale=1000, spk_emb_dim=SPK)
    params = harness_weights(dec)
    dec.load_state_dict(params)
    dec = dec.to(dev).eval()

    B, T, n = args.batch, args.frames, N_STEPS
    z, mask, cond, spk, noise = make_inputs(100 + rank, B, T, n)
    host = [t.pin_memory() for t in (z, mask, cond, spk, noise)]
    zd, md, cd, sd, nd = (t.to(dev) for t in (z, mask, cond, spk, noise))
    from unitspeech_b200.sharding import gather_utterances

    def one_pass():
        out = dec(zd, md, cd, sd, n, text_gradient_scale=TG, spk_gradient_scale=SG, noise=nd)
        if world > 1:
            gather_utterances(out, world * B)      # the job's mels on every rank (NCCL all_gather)
        return out

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 0)):
        out = one_pass()
    sync()
    assert torch.isfinite(out).all(), "non-finite mel from the CUDA decoder"

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = dec.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for _ in range(args.steps):
        one_pass()
    e1.record()
    sync()
    elapsed_ms = e0.elapsed_time(e1)
    launches = dec.launch_count - launches0
    clocks = sampler.finish()
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    frames_total = world * B * T * args.steps
    value = frames_total / (elapsed_ms / 1000.0)

    # ---- end-to-end through the public API with pinned host tensors (copies inside the timed region)
    hz, hm, hc, hs, hn = host
    for _ in range(1):
        dec(hz, hm, hc, hs, n, text_gradient_scale=TG, spk_gradient_scale=SG, noise=hn)
    sync()
    e2e_steps = max(1, min(args.steps, 3))
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(e2e_steps):
        out_h = dec(hz, hm, hc, hs, n, text_gradient_scale=TG, spk_gradient_scale=SG, noise=hn)
    f1.record()
    sync()
    e2e_ms = f0.elapsed_time(f1)
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * T * e2e_steps / (float(t.item()) / 1000.0)
    h2d = sum(x.numel() * 4 for x in host)
    d2h = out_h.numel() * 4

    line = None
    if rank == 0:
        # ---- roofline: one profiled pass of the same workload (per-launch CUDA events on the launching stream)
        peaks = _peaks()
        dec.set_profiling(True)
        dec(zd, md, cd, sd, n, text_gradient_scale=TG, spk_gradient_scale=SG, noise=nd)   # rank-local: no collective here
        torch.cuda.synchronize(dev)
        prof = dec.get_profile()
        dec.set_profiling(False)
        conv_ms, conv_flop, conv_n = prof["conv_igemm"]
        gn_ms, gn_bytes, gn_n = prof["gn_apply"]
        at_ms, at_bytes, at_n = prof["attention"]
        ot_ms, ot_bytes, ot_n = prof["other"]
        achieved = conv_flop / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else None
        total_prof_ms = conv_ms + gn_ms + at_ms + ot_ms
        roofline = {
            "kernel": "conv_igemm_kernel (tcgen05.mma kind::f16, fp16 operands, fp32 TMEM accumulators)",
            "bound": "tensor", "achieved": achieved, "peak": peaks["tensor"], "unit": "TFLOP/s",
            "frac": achieved / peaks["tensor"] if achieved else None, "traffic": None,
            "peak_source": f"{peaks['src']} bf16 GEMM, sustained (burst {peaks['tensor_burst']})",
            "launches_per_pass": conv_n, "ms_per_pass": conv_ms, "share_of_pass": conv_ms / total_prof_ms,
            "flop_per_launch_avg": conv_flop / max(conv_n, 1),
            "frames_per_s_at_peak": peaks["tensor"] * 1e12 / (CONV_FLOP_PER_FRAME_EVAL * 3 * N_STEPS),
        }
        gn_gbs = gn_bytes / (gn_ms / 1e3) / 1e9 if gn_ms > 0 else None
        roofline_hbm = {
            "kernel": "gn_apply_kernel (GroupNorm apply + Mish + embedding/residual + mask, fp16 in/out)",
            "bound": "hbm", "achieved": gn_gbs, "peak": peaks["hbm"], "unit": "GB/s",
            "frac": gn_gbs / peaks["hbm"] if gn_gbs else None, "traffic": None, "launches_per_pass": gn_n,
            "ms_per_pass": gn_ms, "share_of_pass": gn_ms / total_prof_ms,
        }
        breakdown = {"conv_igemm_ms": conv_ms, "gn_apply_ms": gn_ms, "attention_ms": at_ms, "other_ms": ot_ms,
                     "attention_GBps": at_bytes / (at_ms / 1e3) / 1e9 if at_ms > 0 else None,
                     "other_GBps": ot_bytes / (ot_ms / 1e3) / 1e9 if ot_ms > 0 else None}
        # ---- optional second stage (BASELINE.json configs[3]): BigVGAN vocoder on this batch's mels
        # (the two optional legs below report their own failure in the JSON line instead of taking the headline down)
        vocoder = None
        if not args.no_vocoder:
            try:
                vocoder = vocoder_leg(dev, B, T, elapsed_ms / args.steps, peaks)
            except Exception as exc:  # noqa: BLE001
                print(f"vocoder leg failed: {exc!r}", file=sys.stderr)
                vocoder = {"error": repr(exc)}
        # ---- optional: speaker-adaptive fine-tuning step (BASELINE.json configs[4])
        finetune = None
        if not args.no_finetune:
            dec._release()            # frees the sampler's workspace before the training buffers are allocated
            torch.cuda.empty_cache()
            try:
                finetune = finetune_leg(dev, params, peaks, with_cpu=not args.no_cpu and world == 1)
            except Exception as exc:  # noqa: BLE001
                print(f"fine-tune leg failed: {exc!r}", file=sys.stderr)
                finetune = {"error": repr(exc)}
        # ---- CPU baseline on this box's host cores (bounded sample)
        cpu = None
        if not args.no_cpu and world == 1:          # the CPU baseline is an N = 1 figure (torchrun also pins OMP to 1 thread)
            threads = os.cpu_count() or 1
            dt, fps, nst = cpu_sample(params, threads, diffusion_steps=12)   # ~10-20 s of CPU work
            cpu = {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"1 utterance x {T} frames x {nst} of the 50 CFG diffusion steps ({3 * nst} U-Net evaluations, {dt:.1f} s), scaled to the 50-step job"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": f"UnitSpeech decoder (random-init pretrained_decoder architecture, 119.1M params), "
                                   f"50-step text+speaker CFG (1.0/1.0) reverse diffusion, {B} utterances x {T} frames "
                                   f"per GPU (BASELINE.json configs[1]); sharded by utterance across GPUs",
                       "batch_per_gpu": B, "frames": T, "diffusion_steps": n, "cfg_scales": [TG, SG],
                       "estimator_evals_per_step": 3, "parallelism": f"utterance-sharded x{world}",
                       "l2": "per-step working set (multi-GB activations, 131 MB noise) far exceeds the 126 MB L2; no explicit flush",
                       "rtf": (elapsed_ms / 1000.0 / args.steps) / (B * T * 256 / 22050.0)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": launches,
            "roofline": roofline, "roofline_hbm": roofline_hbm, "breakdown_ms_per_pass": breakdown,
            "cpu_baseline": cpu,
            "vocoder_stage": vocoder,
            "finetune_stage": finetune,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        _emit(line)


def finetune_leg(dev, params, peaks, with_cpu=True, B=8, T=176, iters=30):
    """BASELINE.json configs[4]: one fine-tune iteration = zero_grad + diffusion loss forward + backward + clip_grad_norm_ +
    Adam (finetune.py:131-165) on 8 crops of 176 frames (fix_len_compatibility(2 * 22050 // 256)), lr 2e-5, timed with CUDA
    events over `iters` CUDA-graph replays; the oracle's autograd step on the host cores is timed beside it."""
    from unitspeech_b200 import FineTuner, abi
    ft = FineTuner(lr=2e-5, device=dev.index)
    sd = {k: (v * (4.0 * 512) if k.startswith("estimator.final_conv") else v) for k, v in params.items()}   # O(1) score output
    ft.load_state_dict(sd)
    g = torch.Generator().manual_seed(11)
    x0 = (torch.randn(B, N_FEATS, T, generator=g) * 0.5).clamp(-1, 1).to(dev)
    cond = torch.randn(B, N_FEATS, T, generator=g).clamp(-1, 1).to(dev)
    mask = torch.ones(B, 1, T, device=dev)
    spk = torch.randn(B, 1, SPK, generator=g)
    spk = (spk / spk.norm(dim=-1, keepdim=True)).to(dev)
    zs = torch.randn(4, B, N_FEATS, T, generator=g).to(dev)
    ts = torch.rand(4, B, generator=g).clamp(1e-5, 1 - 1e-5).to(dev)
    first = float(ft.train_step(x0, mask, cond, ts[0], spk, zs[0]))
    for i in range(1, 6):
        ft.train_step(x0, mask, cond, ts[i % 4], spk, zs[i % 4])
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ft.train_step(x0, mask, cond, ts[i % 4], spk, zs[i % 4])
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / iters
    last = float(ft.loss)
    assert int(ft.skipped) == 0 and last == last, "fine-tune step produced non-finite gradients"
    # launches per iteration: counted on one eager (non-graph) step, the graph replays the same sequence
    ft.use_cuda_graph = False
    ft._graphs.clear()
    l0 = int(abi.load_library().usb_launch_count(ft.h))
    ft.train_step(x0, mask, cond, ts[0], spk, zs[0])
    torch.cuda.synchronize(dev)
    launches = int(abi.load_library().usb_launch_count(ft.h)) - l0
    flop = 3.0 * CONV_FLOP_PER_FRAME_EVAL * B * T          # forward + data gradient + weight gradient of every conv
    out = {
        "workload": f"fine-tune iteration: {B} crops x {T} frames, U-Net forward + backward of the diffusion loss, "
                    f"clip_grad_norm_(1) + Adam(lr 2e-5) over 119.1M fp32 parameters (BASELINE.json configs[4])",
        "ms_per_iter": ms, "iters_per_s": 1e3 / ms, "seconds_per_500_iters": ms * 0.5, "gpu_launches_per_iter": launches,
        "loss_first": first, "loss_after": last,
        "conv_fwd_bwd": {"bound": "tensor", "achieved": flop / ms / 1e9, "peak": peaks["tensor"], "unit": "TFLOP/s over the whole iteration",
                         "frac": flop / ms / 1e9 / peaks["tensor"]},
    }
    ft.close()
    if with_cpu:
        from oracle import unitspeech_oracle as O
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        Bc = 2
        cpu = lambda v: v[:Bc].cpu()  # noqa: E731
        p_cpu = {k: v.cpu() for k, v in sd.items()}
        t0 = time.perf_counter()
        _, grads = O.loss_t_grads(p_cpu, cpu(x0), cpu(mask), cpu(cond), ts[0][:Bc].cpu(), cpu(spk), zs[0][:Bc].cpu())
        O.clip_and_adam(p_cpu, grads, {}, 1)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 1.0 / (dt * B / Bc), "unit": "iters/s", "cores": threads, "kind": "port",
                               "sample": f"one oracle iteration (autograd + clip + Adam) on {Bc} of the {B} crops ({dt:.1f} s), scaled x{B // Bc}"}
    return out


def vocoder_leg(dev, B, T, decoder_ms, peaks):
    """Times the BigVGAN stage (public 22 kHz / 80-band generator, 112.2M random-init params) on B x T mel frames and
    reports its two kernel classes against their rooflines."""
    from unitspeech_b200 import BigVGAN
    from unitspeech_b200.synthetic import PUBLIC_VOCODER_CONFIG, vocoder_state
    voc = BigVGAN(dict(PUBLIC_VOCODER_CONFIG))
    voc.load_state_dict(vocoder_state(PUBLIC_VOCODER_CONFIG, seed=1))
    voc.to(dev).eval()
    voc.max_frames_per_call = max(8192, T)
    mel = torch.randn(B, N_FEATS, T, device=dev) * 2 - 4
    for _ in range(2):
        wav = voc(mel)
    torch.cuda.synchronize(dev)
    assert torch.isfinite(wav).all(), "non-finite audio from the CUDA vocoder"
    l0 = voc.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 3
    e0.record()
    for _ in range(iters):
        voc(mel)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / iters
    launches = (voc.launch_count - l0) // iters
    voc.set_profiling(True)
    voc(mel)
    torch.cuda.synchronize(dev)
    prof = voc.get_profile()
    voc.set_profiling(False)
    cm, cw, cn = prof["conv_igemm"]
    am, aw, an = prof["snake_act"]
    return {
        "workload": f"BigVGAN 22 kHz/80-band generator (112.2M params, random init), {B} x {T} mel frames -> {B} x {T * 256} samples",
        "ms": ms, "mel_frames_per_s": B * T / ms * 1e3, "audio_seconds_per_s": B * T * 256 / 22050.0 / ms * 1e3,
        "gpu_launches": launches,
        "pipeline_mel_frames_per_s": B * T / (decoder_ms + ms) * 1e3,
        "conv": {"kernel": "conv_igemm (1-D taps, H = 1)", "bound": "tensor", "ms": cm, "launches": cn,
                 "achieved": cw / cm / 1e9, "peak": peaks["tensor"], "unit": "TFLOP/s (channels padded to 64)",
                 "frac": cw / cm / 1e9 / peaks["tensor"]},
        "snake_act": {"kernel": "snake_act_kernel (up x2 + Snake + low-pass + down x2 fused)", "bound": "hbm", "ms": am,
                      "launches": an, "achieved": aw / am / 1e6, "peak": peaks["hbm"], "unit": "GB/s (fp16 read + write)",
                      "frac": aw / am / 1e6 / peaks["hbm"]},
    }


def main():
    # Only the JSON line may reach stdout: libraries (e.g. NCCL's version banner) print there too, so fd 1 is pointed at
    # stderr for the duration of the run and restored for the final print.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        _main()
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if _RESULT_LINE is not None:
        print(_RESULT_LINE, flush=True)


_RESULT_LINE = None


def _emit(line: dict) -> None:
    global _RESULT_LINE
    _RESULT_LINE = json.dumps(line)


def _main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-vocoder", action="store_true", help="skip the vocoder-stage leg")
    ap.add_argument("--no-finetune", action="store_true", help="skip the fine-tune-iteration leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
