#!/usr/bin/env python
"""Benchmark of the reverse-diffusion hot path: mel frames/s of 50-step CFG reverse diffusion.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one full pass of the hot path over one batch: a 50-step text+speaker classifier-free-guided reverse
diffusion (150 U-Net evaluations per utterance) of 32 utterances x 1000 frames per GPU -- the per-GPU shard of
BASELINE.json configs[2] (256 utterances x 1000 frames over 8 GPUs), so `--gpus 8` IS configs[2] and smaller N are the
same shard on fewer GPUs (weak scaling).  The job is `32 N` utterances sharded by utterance through the product's own
`unitspeech_b200.sharding.sample_sharded` (independent sampling loops, no collective in the step loop); the mels are
all-gathered with NCCL inside the timed region.

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the same metric over the same
number of steps through the public API with pinned HOST tensors (H2D of z/cond/mask/spk/noise and D2H of the mels inside
the timed region, gather included); `roofline` = the tcgen05 implicit-GEMM conv kernels (algorithmic conv FLOPs / their
summed CUDA-event time, measured in a separate profiled pass of the same workload, against the measured bf16 GEMM peak;
`traffic` from the committed ncu capture under profiles/); `cpu_baseline` = the unmodified reference (baseline/_ref,
installed by scripts/install_ref.py; the oracle port only if that is absent) on a bounded sample of the same workload;
`gpu_eager_baseline` = the same reference run eagerly by PyTorch/cuDNN on this GPU (fp32 and TF32);
`latency_stage` = one utterance x 256 frames (the reference callers' B = 1 pattern, inference.py:128);
`secondary_16x512` = BASELINE.json configs[1].

`--impl reference` times the reference's own CPU implementation of the path on the host cores: the UNMODIFIED
`UnitSpeech.forward` (unitspeech/unitspeech.py:387-391) imported from baseline/_ref through oracle/ref_shim.py, one
utterance of the same workload per step, with all host threads.
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
BATCH, FRAMES = 32, 1000          # per-GPU shard of BASELINE.json configs[2] (256 x 1000 over 8 GPUs)
N_FEATS, SPK = 80, 256
CONV_FLOP_PER_FRAME_EVAL = 647.27e6   # conv FLOPs per mel frame per estimator evaluation
VOCODER_FLOP_PER_FRAME = 2 * 901859328.0   # BigVGAN public config, FLOPs per mel frame
SCALE = 1.0 / 512                 # harness scale keeping the untrained sampler O(1)


def workload_name(B: int, T: int) -> str:
    return (f"UnitSpeech decoder (random-init pretrained_decoder architecture, 119.1M params), 50-step text+speaker CFG "
            f"(1.0/1.0) reverse diffusion, {B} utterances x {T} frames per GPU"
            + (" = the per-GPU shard of BASELINE.json configs[2] (256 x 1000 over 8 GPUs)" if (B, T) == (32, 1000) else "")
            + "; sharded by utterance across GPUs")


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tensor": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "tensor_burst": float(p["bf16_tflops"]),
                "hbm": float(p["hbm_gbs"]), "src": "measured"}
    return {"tensor": 1400.0, "tensor_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


def _build_info():
    """What the loaded libunitspeech_b200.so was built from (written by unitspeech_b200/build.py next to the library)."""
    try:
        from unitspeech_b200 import build as b
        return b.build_info() or None
    except Exception:  # noqa: BLE001
        return None


def _profile_json(name: str):
    """Committed evidence under profiles/ (ncu-derived traffic, per-step drift) that the line quotes; None if absent."""
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        return None
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:  # noqa: BLE001
        return None


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
# CPU arm: the unmodified reference from baseline/_ref (scripts/install_ref.py), else the oracle port
# ---------------------------------------------------------------------------------------------------------------------
class CpuReference:
    """The reference decoder on the host cores.  kind "reference": the unmodified upstream class
    (unitspeech/unitspeech.py:220) imported from baseline/_ref through oracle/ref_shim.py, called through its own
    `forward` per utterance (B = 1 is its only correct batch) with the per-step randn draws injected;
    kind "port": oracle/unitspeech_oracle.py (only when baseline/_ref is absent)."""

    def __init__(self, params, threads: int, device: str = "cpu"):
        from oracle import ref_shim
        self.threads, self.device = threads, torch.device(device)
        torch.set_num_threads(threads)
        self.params = {k: v.to(self.device) for k, v in params.items()}
        self.shim = ref_shim
        if ref_shim.installed_reference_available():
            U = ref_shim.load_reference(root=ref_shim.INSTALLED_REF)
            dec = U.UnitSpeech(N_FEATS, 128, (1, 2, 4, 8), 0.05, 20, 1000, SPK)
            dec.load_state_dict(params, strict=True)
            self.dec = dec.to(self.device).eval()
            self.kind = "reference"
        else:
            self.dec = None
            self.kind = "port"

    def sample(self, z, mask, cond, spk, noise, n: int, tg: float, sg: float):
        mv = lambda a: a.to(self.device)  # noqa: E731
        with torch.no_grad():
            if self.dec is not None:
                return self.shim.run_reference_per_utterance(self.dec, mv(z), mv(mask), mv(cond), mv(spk), mv(noise), n, tg, sg)
            from oracle import unitspeech_oracle as O
            return O.reverse_diffusion(self.params, mv(z), mv(mask), mv(cond), mv(spk), n, tg, sg, noise=mv(noise))

    def timed(self, T: int, n: int, tg: float, sg: float, seed: int = 7):
        """One utterance x T frames x n diffusion steps through the reference's public call; returns seconds."""
        z, mask, cond, spk, noise = make_inputs(seed, 1, T, n)
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = self.sample(z, mask, cond, spk, noise, n, tg, sg)
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert torch.isfinite(out).all()
        return dt


REF_DIFFUSION_STEPS = 2     # diffusion steps per reference-arm bench step (n_timesteps = 2 is the smallest call the reference supports)


def reference_sample_text(T: int, n: int) -> str:
    return (f"1 utterance x {T} frames x n_timesteps={n} text+speaker CFG diffusion steps ({3 * n} U-Net evaluations) through the "
            f"reference's UnitSpeech.forward per bench step; per-step cost is independent of n_timesteps, so "
            f"frames/s of the 50-step job = {T} / (50 x seconds per diffusion step)")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    params = harness_weights()
    ref = CpuReference(params, threads)
    T, n = args.frames, REF_DIFFUSION_STEPS
    for _ in range(min(max(args.warmup, 0), 1)):      # one warm-up call is enough for a CPU arm (thread pool, allocator)
        ref.timed(T, n, TG, SG)
    per_step = []
    t_total = 0.0
    for i in range(args.steps):
        dt = ref.timed(T, n, TG, SG, seed=7 + i)
        per_step.append(dt / n)
        t_total += dt
    v = T / (N_STEPS * statistics.median(per_step))
    sample = reference_sample_text(T, n)
    cfg1 = None
    if not args.no_config1:
        # BASELINE.md section 3 / SURVEY 8(d6): BASELINE.json configs[0] exactly -- 1 utterance x 256 frames, 50 steps, no CFG,
        # the whole call, best of 3
        best = min(ref.timed(256, N_STEPS, 0.0, 0.0, seed=3) for _ in range(3))
        cfg1 = {"workload": "BASELINE.json configs[0]: 1 utterance x 80 mel x 256 frames, 50 steps, no CFG, fp32 CPU, whole call",
                "seconds_best_of_3": best, "value": 256 / best, "unit": UNIT, "rtf": best / (256 * 256 / 22050.0)}
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * t_total / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.batch, T), "batch_per_gpu": args.batch, "frames": T,
                   "diffusion_steps": N_STEPS, "cfg_scales": [TG, SG], "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": ref.kind, "sample": sample,
                         "install": "baseline/_ref (unmodified reference files, scripts/install_ref.py)" if ref.kind == "reference"
                         else "oracle/unitspeech_oracle.py (baseline/_ref absent)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "config1_cpu": cfg1,
        "seconds_per_diffusion_step": {"min": min(per_step), "median": statistics.median(per_step), "max": max(per_step)},
        "gpu_launches": 0,
    }
    _emit(line)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def _new_decoder(dev, params=None):
    from unitspeech_b200 import UnitSpeech
    dec = UnitSpeech(N_FEATS, 128, (1, 2, 4, 8), beta_min=0.05, beta_max=20, pe_scale=1000, spk_emb_dim=SPK)
    params = harness_weights(dec) if params is None else params
    dec.load_state_dict(params)
    return dec.to(dev).eval(), params


def _time_passes(fn, steps, sync):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    sync()
    return e0.elapsed_time(e1), out


def run_gpu_arm(args):
    import torch.distributed as dist
    from unitspeech_b200.sharding import sample_sharded, shard_range

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

    dec, params = _new_decoder(dev)
    B, T, n = args.batch, args.frames, N_STEPS
    n_job = world * B                               # utterances of the whole job; this rank samples shard_range(n_job, rank, world)
    b0, b1 = shard_range(n_job, rank, world)
    assert b1 - b0 == B
    # the job's small tensors exist in full on every rank (as a caller of sample_sharded holds them); the per-step noise
    # (50x larger) only for the rank's own utterances -- sample_sharded(noise_is_local=True)
    zj, mj, cj, sj, _ = make_inputs(100, n_job, T, 1)
    g = torch.Generator().manual_seed(1000 + rank)
    noise = torch.randn(n, B, N_FEATS, T, generator=g) * SCALE
    host = [t.pin_memory() for t in (zj, mj, cj, sj, noise)]
    zd, md, cd, sd, nd = (t.to(dev) for t in host)

    def one_pass(tensors=(zd, md, cd, sd, nd)):
        z_, m_, c_, s_, n_ = tensors
        return sample_sharded(dec, z_, m_, c_, s_, n, TG, SG, noise=n_, noise_is_local=True)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 0)):
        out = one_pass()
    sync()
    assert out.shape[0] == n_job and torch.isfinite(out).all(), "non-finite mel from the CUDA decoder"
    sat = dec.saturation_count()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = dec.launch_count
    elapsed_ms, _ = _time_passes(one_pass, args.steps, sync)
    launches = dec.launch_count - launches0
    clocks = sampler.finish()
    elapsed_ms = max_over_ranks(elapsed_ms)
    frames_total = n_job * T * args.steps
    value = frames_total / (elapsed_ms / 1000.0)

    # ---- end-to-end through the public API with pinned host tensors: H2D of the rank's inputs, D2H of its mels and the
    # gather of the job's mels inside the timed region, over the same number of steps as `value`
    one_pass(host)
    e2e_ms, out_h = _time_passes(lambda: one_pass(host), args.steps, sync)
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_value = frames_total / (e2e_ms / 1000.0)
    h2d = sum(x[b0:b1].numel() * 4 for x in host[:4]) + host[4].numel() * 4     # per rank and step
    d2h = B * N_FEATS * T * 4

    line = None
    if rank == 0:
        # ---- roofline: one profiled pass of the same workload (per-launch CUDA events on the launching stream)
        peaks = _peaks()
        dec.set_profiling(True)
        dec(zd[b0:b1], md[b0:b1], cd[b0:b1], sd[b0:b1], n, text_gradient_scale=TG, spk_gradient_scale=SG, noise=nd)   # rank-local
        torch.cuda.synchronize(dev)
        prof = dec.get_profile()
        dec.set_profiling(False)
        conv_ms, conv_flop, conv_n = prof["conv_igemm"]
        gn_ms, gn_bytes, gn_n = prof["gn_apply"]
        at_ms, at_bytes, at_n = prof["attention"]
        ot_ms, ot_bytes, ot_n = prof["other"]
        achieved = conv_flop / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else None
        total_prof_ms = conv_ms + gn_ms + at_ms + ot_ms
        traffic = _profile_json("r2_traffic.json") or {}
        conv_traffic = (traffic.get(f"{B}x{T}") or {}).get("conv_igemm")
        roofline = {
            "kernel": "conv_igemm_{halo,swapped,}_kernel (tcgen05.mma kind::f16, fp16 operands, fp32 TMEM accumulators), all 60 conv launches of an evaluation",
            "bound": "tensor", "achieved": achieved, "peak": peaks["tensor"], "unit": "TFLOP/s",
            "frac": achieved / peaks["tensor"] if achieved else None,
            "traffic": conv_traffic["dram_bytes_per_launch"] if conv_traffic else None,
            "traffic_note": (conv_traffic or {}).get("note"),
            "peak_source": f"{peaks['src']} bf16 GEMM, sustained (burst {peaks['tensor_burst']})",
            "launches_per_pass": conv_n, "ms_per_pass": conv_ms, "share_of_pass": conv_ms / total_prof_ms,
            "flop_per_launch_avg": conv_flop / max(conv_n, 1),
            "algorithmic_bytes_per_launch_avg": (conv_traffic or {}).get("algorithmic_bytes_per_launch"),
            "whole_step_frac": CONV_FLOP_PER_FRAME_EVAL * 3 * N_STEPS * (value / world) / 1e12 / peaks["tensor"],
            "frames_per_s_at_peak": peaks["tensor"] * 1e12 / (CONV_FLOP_PER_FRAME_EVAL * 3 * N_STEPS),
        }
        gn_gbs = gn_bytes / (gn_ms / 1e3) / 1e9 if gn_ms > 0 else None
        gn_traffic = (traffic.get(f"{B}x{T}") or {}).get("gn_apply")
        roofline_hbm = {
            "kernel": "gn_apply_kernel (GroupNorm apply + Mish + embedding/residual + mask, fp16 in/out)",
            "bound": "hbm", "achieved": gn_gbs, "peak": peaks["hbm"], "unit": "GB/s",
            "frac": gn_gbs / peaks["hbm"] if gn_gbs else None,
            "traffic": gn_traffic["dram_bytes_per_launch"] if gn_traffic else None, "launches_per_pass": gn_n,
            "ms_per_pass": gn_ms, "share_of_pass": gn_ms / total_prof_ms,
        }
        breakdown = {"conv_igemm_ms": conv_ms, "gn_apply_ms": gn_ms, "attention_ms": at_ms, "other_ms": ot_ms,
                     "attention_GBps": at_bytes / (at_ms / 1e3) / 1e9 if at_ms > 0 else None,
                     "other_GBps": ot_bytes / (ot_ms / 1e3) / 1e9 if ot_ms > 0 else None}
        legs = {}

        def leg(name, fn):
            # the optional legs report their own failure in the JSON line instead of taking the headline down
            try:
                legs[name] = fn()
            except Exception as exc:  # noqa: BLE001
                print(f"{name} leg failed: {exc!r}", file=sys.stderr)
                legs[name] = {"error": repr(exc)}

        if not args.no_secondary:
            leg("secondary_16x512", lambda: small_config_leg(dec, dev, 16, 512, 3, peaks, "BASELINE.json configs[1]"))
            leg("latency_stage", lambda: small_config_leg(dec, dev, 1, 256, 5, peaks,
                                                          "one utterance x 256 frames (2.97 s of speech), the reference callers' "
                                                          "B = 1 pattern (inference.py:128)"))
        if not args.no_vocoder:
            leg("vocoder_stage", lambda: vocoder_leg(dev, 16, 512, None, peaks))
        dec._release()            # frees the sampler's workspace before the training / eager buffers are allocated
        torch.cuda.empty_cache()
        if not args.no_finetune:
            leg("finetune_stage", lambda: finetune_leg(dev, params, peaks, with_cpu=not args.no_cpu and world == 1))
        if not args.no_eager and world == 1:
            leg("gpu_eager_baseline", lambda: eager_leg(dev, params, T))
        # ---- CPU baseline on this box's host cores (bounded sample; an N = 1 figure: torchrun pins OMP to 1 thread)
        cpu = None
        if not args.no_cpu and world == 1:
            threads = os.cpu_count() or 1
            ref = CpuReference(params, threads)
            ref.timed(T, REF_DIFFUSION_STEPS, TG, SG)
            dts = [ref.timed(T, REF_DIFFUSION_STEPS, TG, SG, seed=8 + i) for i in range(3)]      # ~10-20 s of CPU work
            fps = T / (N_STEPS * statistics.median(dts) / REF_DIFFUSION_STEPS)
            cpu = {"value": fps, "unit": UNIT, "cores": threads, "kind": ref.kind,
                   "sample": reference_sample_text(T, REF_DIFFUSION_STEPS) + f"; median of 3 calls ({sum(dts):.1f} s)"}
        drift = _profile_json(f"r2_drift_T{T}.json")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": workload_name(B, T),
                       "batch_per_gpu": B, "frames": T, "job_utterances": n_job, "diffusion_steps": n, "cfg_scales": [TG, SG],
                       "estimator_evals_per_step": 3, "parallelism": f"utterance-sharded x{world} (sample_sharded)",
                       "l2": "per-step working set (tens of GB of activations, 0.5 GB of noise) far exceeds the 126 MB L2; no explicit flush",
                       "rtf": (elapsed_ms / 1000.0 / args.steps) / (n_job * T * 256 / 22050.0)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": args.steps, "ms_per_step": e2e_ms / args.steps, "includes_gather": world > 1},
            "gpu_launches": launches,
            "fp16_saturation_events": sat,
            "roofline": roofline, "roofline_hbm": roofline_hbm, "breakdown_ms_per_pass": breakdown,
            "cpu_baseline": cpu,
            "per_step_drift": ({"file": f"profiles/r2_drift_T{T}.json", **{k: drift[k] for k in drift if k != "per_step"}}
                               if drift else None),
            "build": _build_info(),
        }
        line.update(legs)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        _emit(line)


def small_config_leg(dec, dev, B, T, passes, peaks, what):
    """A second workload on the same handle: ms per 50-step CFG pass (CUDA events, inputs resident) and the conv class
    against the tensor roofline from one profiled pass."""
    n = N_STEPS
    z, mask, cond, spk, noise = (t.to(dev) for t in make_inputs(5, B, T, n))
    run = lambda: dec(z, mask, cond, spk, n, text_gradient_scale=TG, spk_gradient_scale=SG, noise=noise)  # noqa: E731
    for _ in range(2):
        run()
    l0 = dec.launch_count
    ms, _ = _time_passes(run, passes, lambda: torch.cuda.synchronize(dev))
    launches = (dec.launch_count - l0) // passes
    ms /= passes
    dec.set_profiling(True)
    run()
    torch.cuda.synchronize(dev)
    prof = dec.get_profile()
    dec.set_profiling(False)
    conv_ms, conv_flop, _ = prof["conv_igemm"]
    # the profiled pass serialises on per-launch events; the conv fraction below uses the un-profiled pass time
    flop = CONV_FLOP_PER_FRAME_EVAL * 3 * n * B * T
    return {"workload": f"{B} x {T} frames, 50-step text+speaker CFG: {what}", "ms_per_pass": ms,
            "frames_per_s": B * T / ms * 1e3, "rtf": ms / 1e3 / (B * T * 256 / 22050.0), "gpu_launches_per_pass": launches,
            "whole_pass_TFLOPs": flop / ms / 1e9, "whole_pass_frac": flop / ms / 1e9 / peaks["tensor"],
            "conv_class_TFLOPs_profiled": conv_flop / conv_ms / 1e9 if conv_ms > 0 else None,
            "conv_class_frac_profiled": conv_flop / conv_ms / 1e9 / peaks["tensor"] if conv_ms > 0 else None}


def eager_leg(dev, params, T):
    """BASELINE.md section 3 'second baseline': the reference modules run EAGERLY by PyTorch/cuDNN on this same B200
    (per-utterance B = 1, the only batch the reference supports), fp32 (cudnn.allow_tf32 = False, the parity oracle's
    arithmetic) and TF32 (PyTorch's default): what a user of the reference gets on this GPU today."""
    ref = CpuReference(params, os.cpu_count() or 1, device=str(dev))
    out = {"what": "unmodified reference (baseline/_ref) eager on this GPU" if ref.kind == "reference" else "oracle port eager on this GPU",
           "workload": f"1 utterance x {T} frames, 50-step text+speaker CFG (1.0/1.0), whole UnitSpeech.forward call", "unit": UNIT}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            ref.timed(T, 2, TG, SG)
            dt = ref.timed(T, N_STEPS, TG, SG)
            out[f"frames_per_s_{name}"] = T / dt
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    del ref
    torch.cuda.empty_cache()
    return out


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
        "pipeline_mel_frames_per_s": B * T / (decoder_ms + ms) * 1e3 if decoder_ms else None,
        "conv": {"kernel": "conv_igemm (1-D taps, H = 1)", "bound": "tensor", "ms": cm, "launches": cn,
                 # algorithmic FLOPs: SURVEY 8(d4) 1.804 GFLOP per mel frame (901 859 328 MAC), not the padded channel counts
                 "achieved": VOCODER_FLOP_PER_FRAME * B * T / cm / 1e9, "peak": peaks["tensor"], "unit": "TFLOP/s (algorithmic)",
                 "frac": VOCODER_FLOP_PER_FRAME * B * T / cm / 1e9 / peaks["tensor"],
                 "padded_TFLOPs": cw / cm / 1e9},
        "whole_stage_frac": VOCODER_FLOP_PER_FRAME * B * T / ms / 1e9 / peaks["tensor"],
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
    ap.add_argument("--no-eager", action="store_true", help="skip the eager-PyTorch-on-this-GPU baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the 16 x 512 and 1 x 256 legs")
    ap.add_argument("--no-config1", action="store_true", help="reference arm: skip the BASELINE configs[0] whole-call timing")
    ap.add_argument("--headline-only", action="store_true", help="only value / e2e / roofline (no optional legs, no CPU baseline)")
    args = ap.parse_args()
    if args.headline_only:
        args.no_cpu = args.no_vocoder = args.no_finetune = args.no_eager = args.no_secondary = True
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
