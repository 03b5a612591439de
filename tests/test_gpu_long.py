"""Parity on the configurations bench.py MEASURES ("parity is green but not on the measured configs"):

  * one utterance x 512 and x 1000 frames, 50 steps, text+speaker CFG 1.0/1.0, full-size network, ragged length, against
    `oracle.reverse_diffusion` on the CPU -- the per-step drift (max-abs / mean-abs of x_t after every step) is written
    to gpurun_out/r2_drift_T{T}.json and kept under profiles/ (north_star: "with per-step drift reported");
  * the BigVGAN public configuration at T = 512 and 1000 frames against `oracle.bigvgan_oracle`;
  * fine-tune gradients and one clip + Adam step at BASELINE.json configs[4]'s size, 8 crops x 176 frames, full network,
    against `oracle.loss_t_grads` / `clip_and_adam`.

Bitwise batch invariance (tests/test_gpu_decoder.py) makes the batch size irrelevant for the sampler, so one utterance
pins what the long sequence adds: GroupNorm fixed-point sums over 1.28 M elements, the attention softmax over 80 000
positions, 50 steps of accumulated fp16 rounding.

Tolerances (BASELINE.json north_star): final mel max-abs <= 1e-2, mean-abs <= 1e-3 in normalised mel space."""

import json
import os

import pytest
import torch

from oracle import bigvgan_oracle as V
from oracle import unitspeech_oracle as O

pytestmark = pytest.mark.gpu

MAX_TOL, MEAN_TOL = 1e-2, 1e-3
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _out_dir():
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    return d


@pytest.mark.parametrize("T,length", [(512, 487), (1000, 937)])
def test_fifty_step_cfg_parity_and_drift_at_bench_length(T, length):
    from unitspeech_b200 import UnitSpeech
    torch.set_num_threads(os.cpu_count() or 1)
    n, s = 50, 1.0 / 512
    p = O.harness_params(seed=1234, out_scale=s)
    z, mask, cond, spk, noise = O.harness_inputs(1, T, n, seed=21, scale=s, lengths=(length,))
    ref_trace = []
    ref = O.reverse_diffusion(p, z, mask, cond, spk, n, 1.0, 1.0, noise=noise, trace=ref_trace)
    dec = UnitSpeech(80, 128, (1, 2, 4, 8), spk_emb_dim=256)
    dec.load_state_dict(p, strict=True)
    dec = dec.cuda().eval()
    out, tr = dec.reverse_diffusion(z.cuda(), mask.cuda(), cond.cuda(), spk.cuda(), n, 1.0, 1.0, noise=noise.cuda(), trace=True)
    tr = tr.cpu()
    d = (out.cpu() - ref).abs()
    per_step = [{"step": i, "max_abs": float((tr[i] - ref_trace[i]).abs().max()),
                 "mean_abs": float((tr[i] - ref_trace[i]).abs().mean()),
                 "ref_abs_max": float(ref_trace[i].abs().max())} for i in range(n)]
    rec = {"what": "per-step drift of x_t, CUDA decoder vs fp32 CPU oracle (oracle/unitspeech_oracle.py), same weights, inputs and injected noise",
           "utterances": 1, "frames": T, "length": length, "steps": n, "cfg_scales": [1.0, 1.0],
           "final_max_abs": float(d.max()), "final_mean_abs": float(d.mean()), "final_ref_abs_max": float(ref.abs().max()),
           "worst_step_max_abs": max(r["max_abs"] for r in per_step), "worst_step_mean_abs": max(r["mean_abs"] for r in per_step),
           "tolerance": {"max_abs": MAX_TOL, "mean_abs": MEAN_TOL}, "fp16_saturation_events": dec.saturation_count(),
           "device": torch.cuda.get_device_name(0), "per_step": per_step}
    with open(os.path.join(_out_dir(), f"r2_drift_T{T}.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(f"T={T}: final max-abs {rec['final_max_abs']:.3e} mean-abs {rec['final_mean_abs']:.3e} |ref|max {rec['final_ref_abs_max']:.3f}; "
          f"worst step max-abs {rec['worst_step_max_abs']:.3e}")
    assert 0.5 < float(ref.abs().max()) < 8.0          # normalised-mel regime
    assert rec["final_max_abs"] <= MAX_TOL and rec["final_mean_abs"] <= MEAN_TOL
    assert float(out[0, :, length:].abs().max()) == 0.0
    assert rec["fp16_saturation_events"] == 0


@pytest.mark.parametrize("T", [512, 1000])
def test_vocoder_public_config_at_bench_length(T):
    from unitspeech_b200 import BigVGAN
    torch.set_num_threads(os.cpu_count() or 1)
    h = dict(V.PUBLIC_22KHZ_80BAND)
    mel = torch.randn(1, 80, T, generator=torch.Generator().manual_seed(31)) * 2 - 4
    params = V.harness_params(h)
    ref = V.bigvgan_forward(params, mel, h)
    voc = BigVGAN(h)
    voc.load_state_dict(params)
    voc = voc.cuda().eval()
    out = voc(mel.cuda()).cpu()
    assert out.shape == (1, 1, T * 256)
    d = (out - ref).abs()
    print(f"vocoder T={T}: max-abs {float(d.max()):.3e} mean-abs {float(d.mean()):.3e} ref-absmax {float(ref.abs().max()):.3f}")
    assert float(d.max()) <= MAX_TOL and float(d.mean()) <= MEAN_TOL


def test_finetune_gradients_and_adam_step_at_config5_size():
    """8 crops x 176 frames (fix_len_compatibility(2 * 22050 // 256)), full network: every parameter gradient vs the
    oracle's fp32 autograd, then one clip_grad_norm_(1) + Adam(lr 2e-5) step vs `clip_and_adam`."""
    from unitspeech_b200 import FineTuner
    torch.set_num_threads(os.cpu_count() or 1)
    B, T, s, lr = 8, 176, 4.0, 2e-5
    params = O.harness_params(seed=1234, out_scale=s)
    lengths = (176, 176, 150, 176, 101, 176, 176, 64)
    _, mask, cond, spk, _ = O.harness_inputs(B, T, 2, seed=41, lengths=lengths)
    g = torch.Generator().manual_seed(42)
    x0 = (torch.randn(B, 80, T, generator=g) * 0.5).clamp(-1, 1) * mask
    t = torch.rand(B, generator=g).clamp(1e-5, 1 - 1e-5)
    z = torch.randn(B, 80, T, generator=g)
    ref_loss, ref = O.loss_t_grads(params, x0, mask, cond, t, spk, z)
    ft = FineTuner(lr=lr, max_norm=1.0)
    ft.load_state_dict(params)
    ft.use_cuda_graph = False
    ft.zero_grad()
    loss = float(ft.forward(x0, mask, cond, t, spk, z))
    ft.backward()
    torch.cuda.synchronize()
    assert loss == pytest.approx(float(ref_loss), rel=5e-3)
    got = {k: v.cpu() for k, v in ft.unscaled_grads().items()}
    scalars = [k for k, r in ref.items() if r.numel() == 1 and k.endswith(".fn.g")]
    g_scale = float(torch.stack([ref[k].reshape(()) for k in scalars]).norm())
    worst = (0.0, "")
    for k, r in ref.items():
        gk = got[k]
        nr = float(r.norm())
        if nr == 0.0:
            assert float(gk.abs().max()) == 0.0, k
            continue
        if k in scalars:
            assert abs(float(gk) - float(r)) <= 2e-2 * g_scale, k
            continue
        rel = float((gk - r).norm()) / nr
        cos = float(torch.dot(gk.reshape(-1), r.reshape(-1)) / (gk.norm() * r.norm()))
        worst = max(worst, (rel, k))
        assert rel <= 2e-2 and cos >= 0.999, f"{k}: rel {rel:.3e} cos {cos:.5f}"
    total = float(torch.sqrt(sum((v.double() ** 2).sum() for v in ref.values())))
    # one optimizer step from the same state
    loss2 = float(ft.train_step(x0, mask, cond, t, spk, z))
    assert loss2 == pytest.approx(loss, rel=1e-6)
    assert ft.grad_norm() == pytest.approx(total, rel=2e-2)
    assert int(ft.skipped) == 0
    p_or, _ = O.clip_and_adam(dict(params), ref, {}, 1, lr=lr)
    new = {k: v.cpu() for k, v in ft.state_dict().items()}
    bad = []
    for k in params:
        d, do = new[k] - params[k], p_or[k] - params[k]
        if float(do.abs().max()) == 0.0:
            assert float(d.abs().max()) == 0.0, k
            continue
        # first Adam step: every element moves by ~lr * sign(g); compare distance and direction
        if not (float(d.norm()) == pytest.approx(float(do.norm()), rel=5e-2)
                and float(torch.dot(d.reshape(-1), do.reshape(-1)) / (d.norm() * do.norm())) >= 0.95):
            bad.append(k)
    print(f"fine-tune 8 x 176: loss {loss:.6f} (oracle {float(ref_loss):.6f}); |g| {ft.grad_norm():.4e} (oracle {total:.4e}); "
          f"worst gradient rel error {worst[0]:.2e} ({worst[1]})")
    assert not bad, bad[:5]
    ft.close()
