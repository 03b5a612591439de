"""End-to-end parity of the CUDA decoder (through the UnitSpeech-compatible class -> C ABI) against the CPU oracle
and the committed reference-generated golden vectors.

Tolerance (BASELINE.json north_star): final mel max-abs <= 1e-2 and mean-abs <= 1e-3 in normalised mel space
(|x| = O(1)).  For vectors whose scale is not O(1) the same bounds are applied relative to max(1, |ref|_inf).
"""

import os
import sys

import numpy as np
import pytest
import torch

from oracle import unitspeech_oracle as O

pytestmark = pytest.mark.gpu

MAX_TOL, MEAN_TOL = 1e-2, 1e-3


def _decoder(dim, mults, params):
    from unitspeech_b200 import UnitSpeech
    dec = UnitSpeech(n_feats=80, dim=dim, dim_mults=mults, beta_min=0.05, beta_max=20, pe_scale=1000, spk_emb_dim=256)
    dec.load_state_dict(params, strict=True)
    return dec.cuda().eval()


def _errs(out, ref):
    d = (out.detach().cpu().float() - ref).abs()
    scale = max(1.0, float(ref.abs().max()))
    return float(d.max()) / scale, float(d.mean()) / scale


def _load(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    dim, B, T, n = (int(v) for v in g["meta"])
    mults = tuple(int(v) for v in g["mults"])
    lengths = tuple(int(v) for v in g["lengths"])
    tg, sg, s = (float(v) for v in g["scales"])
    return g, dim, mults, B, T, n, lengths, tg, sg, s


GOLDEN = ["d64_cfg", "d64_nocfg", "d64_textonly", "d64_spkonly", "full_cfg", "full_nocfg10"]


@pytest.mark.parametrize("name", GOLDEN)
def test_reverse_diffusion_matches_reference_golden(golden_dir, name):
    g, dim, mults, B, T, n, lengths, tg, sg, s = _load(golden_dir, name)
    p = O.harness_params(dim=dim, dim_mults=mults, seed=1234, out_scale=s)
    z, mask, cond, spk, noise = O.harness_inputs(B, T, n, seed=11, scale=s, lengths=lengths)
    dec = _decoder(dim, mults, p)
    out = dec(z.cuda(), mask.cuda(), cond.cuda(), spk.cuda(), n, text_gradient_scale=tg, spk_gradient_scale=sg,
              noise=noise.cuda())
    assert out.shape == (B, 80, T) and out.is_cuda
    mx, mn = _errs(out, torch.from_numpy(g["out"]))
    print(f"{name}: rel max-abs {mx:.3e} mean-abs {mn:.3e}")
    assert mx <= MAX_TOL and mn <= MEAN_TOL


@pytest.mark.parametrize("name", GOLDEN)
def test_estimator_matches_reference_golden(golden_dir, name):
    g, dim, mults, B, T, n, lengths, tg, sg, s = _load(golden_dir, name)
    p = O.harness_params(dim=dim, dim_mults=mults, seed=1234, out_scale=s)
    z, mask, cond, spk, noise = O.harness_inputs(B, T, n, seed=11, scale=s, lengths=lengths)
    dec = _decoder(dim, mults, p)
    est = dec.estimator(z.cuda(), mask.cuda(), cond.cuda(), torch.full((B,), 0.37).cuda(), spk.cuda())
    ref = torch.from_numpy(g["est"])
    d = (est.cpu() - ref).abs()
    scale = float(ref.abs().max())
    print(f"{name}: estimator rel max {float(d.max()) / scale:.3e} mean {float(d.mean()) / scale:.3e}")
    # one evaluation, fp16 operands/activations: tf32-class accuracy relative to the output scale
    assert float(d.max()) <= 1e-2 * scale and float(d.mean()) <= 2e-3 * scale


def test_fifty_step_cfg_parity_and_drift():
    """The headline setting: 50 steps, text+speaker CFG 1.0/1.0, full-size network, ragged batch; oracle on CPU."""
    B, T, n, s = 2, 64, 50, 1.0 / 512
    p = O.harness_params(seed=1234, out_scale=s)
    z, mask, cond, spk, noise = O.harness_inputs(B, T, n, seed=3, scale=s, lengths=(64, 50))
    ref_trace = []
    ref = O.reverse_diffusion(p, z, mask, cond, spk, n, 1.0, 1.0, noise=noise, trace=ref_trace)
    dec = _decoder(128, (1, 2, 4, 8), p)
    out, tr = dec.reverse_diffusion(z.cuda(), mask.cuda(), cond.cuda(), spk.cuda(), n, 1.0, 1.0, noise=noise.cuda(),
                                    trace=True)
    d = (out.cpu() - ref).abs()
    drift = [float((tr[i].cpu() - ref_trace[i]).abs().max()) for i in range(n)]
    print(f"final |ref|max {float(ref.abs().max()):.3f} max-abs {float(d.max()):.3e} mean-abs {float(d.mean()):.3e}")
    print("per-step drift (max-abs) every 5 steps:", ["%.2e" % v for v in drift[::5]])
    assert 0.5 < float(ref.abs().max()) < 8.0          # normalised-mel regime
    assert float(d.max()) <= MAX_TOL and float(d.mean()) <= MEAN_TOL
    # padded frames stay exactly zero
    assert float(out[1, :, 50:].abs().max()) == 0.0


def test_batch_equals_per_utterance_calls():
    """Batched call == the same utterances sampled one by one (the reference's only correct mode)."""
    B, T, n, s = 3, 32, 4, 1.0 / 512
    p = O.harness_params(seed=1234, out_scale=s)
    z, mask, cond, spk, noise = O.harness_inputs(B, T, n, seed=5, scale=s, lengths=(32, 17, 24))
    dec = _decoder(128, (1, 2, 4, 8), p)
    args = lambda sl: (z[sl].cuda(), mask[sl].cuda(), cond[sl].cuda(), spk[sl].cuda())  # noqa: E731
    full = dec(*args(slice(0, B)), n, 1.0, 1.0, noise=noise.cuda())
    for b in range(B):
        one = dec(*args(slice(b, b + 1)), n, 1.0, 1.0, noise=noise[:, b:b + 1].cuda())
        # identical arithmetic per utterance: tiles never mix samples and the GroupNorm partial sums are
        # accumulated with integer (associative) atomics -> bitwise batch invariance
        assert torch.equal(one, full[b:b + 1])


def test_graph_replayed_sampler_is_bit_identical_to_the_eager_loop():
    """Small calls replay ONE captured CUDA graph per diffusion step (device step counter + scalar table)."""
    B, T, n, s = 2, 32, 6, 1.0 / 512
    p = O.harness_params(seed=1234, out_scale=s)
    z, mask, cond, spk, noise = O.harness_inputs(B, T, n, seed=5, scale=s, lengths=(32, 21))
    dec = _decoder(128, (1, 2, 4, 8), p)
    args = (z.cuda(), mask.cuda(), cond.cuda(), spk.cuda(), n, 1.0, 1.0)
    dec.graph_mode = 0
    eager = dec(*args, noise=noise.cuda())
    assert dec.graph_steps == 0
    dec.graph_mode = -1                       # auto: 6 rows x 32 frames is far below the threshold
    l0 = dec.launch_count
    g1 = dec(*args, noise=noise.cuda())
    assert dec.graph_steps == n - 1           # step 0 runs eagerly, the rest are replays
    g2 = dec(*args, noise=noise.cuda())       # second call reuses the instantiated graph
    assert dec.graph_steps == 2 * (n - 1)
    assert torch.equal(g1, eager) and torch.equal(g2, eager)
    assert dec.launch_count - l0 > 2 * n * 100      # replayed kernels are counted
    # other guidance scales / branch sets / step counts / inputs through the same handle
    for tg, sg, steps in ((0.5, 2.0, 6), (1.0, 0.0, 4), (0.0, 0.0, 9)):
        nz = torch.randn(steps, B, 80, T, generator=torch.Generator().manual_seed(steps)).cuda() * s
        dec.graph_mode = 0
        want = dec(z.cuda(), mask.cuda(), cond.cuda() * 0.5, spk.cuda(), steps, tg, sg, noise=nz)
        dec.graph_mode = 1
        got = dec(z.cuda(), mask.cuda(), cond.cuda() * 0.5, spk.cuda(), steps, tg, sg, noise=nz)
        assert torch.equal(got, want), (tg, sg, steps)
    # host-buffer entry and noise=None also go through the replayed step
    dec.graph_mode = 1
    assert torch.equal(dec(z, mask, cond, spk, n, 1.0, 1.0, noise=noise), eager.cpu())


def test_graph_mode_survives_shape_and_step_count_changes():
    """The captured step belongs to one (rows, frames) plan: changing the shape re-plans and re-captures; the step count,
    the noise and the inputs are not part of the graph."""
    s = 1.0 / 32
    p = O.harness_params(dim=64, dim_mults=(1, 2), seed=1234, out_scale=s)
    dec = _decoder(64, (1, 2), p)
    want = {}
    for mode in (0, 1):
        dec.graph_mode = mode
        for B, T, n, seed in ((1, 32, 5, 1), (2, 16, 4, 2), (1, 32, 7, 3), (2, 16, 3, 4), (1, 8, 4, 5)):
            z, mask, cond, spk, noise = O.harness_inputs(B, T, n, seed=seed, scale=s)
            out = dec(z.cuda(), mask.cuda(), cond.cuda(), spk.cuda(), n, 1.0, 1.0, noise=noise.cuda())
            if mode == 0:
                want[(B, T, n, seed)] = out
            else:
                assert torch.equal(out, want[(B, T, n, seed)]), (B, T, n)
    assert dec.graph_steps > 0


def test_split_k_mode_is_deterministic_batch_invariant_and_within_tolerance():
    """Latency mode: the level-2/3 convolutions split K over idle SMs; partial tiles are summed in split order."""
    B, T, n, s = 3, 64, 4, 1.0 / 512
    p = O.harness_params(seed=1234, out_scale=s)
    z, mask, cond, spk, noise = O.harness_inputs(B, T, n, seed=6, scale=s, lengths=(64, 40, 57))
    ref = O.reverse_diffusion(p, z, mask, cond, spk, n, 1.0, 1.0, noise=noise)
    dec = _decoder(128, (1, 2, 4, 8), p)
    args = lambda sl: (z[sl].cuda(), mask[sl].cuda(), cond[sl].cuda(), spk[sl].cuda())  # noqa: E731
    outs = {}
    for mode in (0, 1):
        dec.splitk_mode = mode
        full = dec(*args(slice(0, B)), n, 1.0, 1.0, noise=noise.cuda())
        assert torch.equal(dec(*args(slice(0, B)), n, 1.0, 1.0, noise=noise.cuda()), full)        # run-to-run
        for b in range(B):                                                                         # batch invariance
            one = dec(*args(slice(b, b + 1)), n, 1.0, 1.0, noise=noise[:, b:b + 1].cuda())
            assert torch.equal(one, full[b:b + 1]), (mode, b)
        mx, mn = _errs(full, ref)
        print(f"split-K mode {mode}: rel max-abs {mx:.3e} mean-abs {mn:.3e}")
        assert mx <= MAX_TOL and mn <= MEAN_TOL
        outs[mode] = full
    # the two modes round differently (documented); both are within tolerance of the fp32 oracle
    assert float((outs[0] - outs[1]).abs().max()) <= 2e-3
    with pytest.raises(ValueError):
        dec.splitk_mode = 2


def test_noise_none_follows_reference_rng_order():
    B, T, n, s = 1, 16, 3, 1.0 / 32
    p = O.harness_params(dim=64, dim_mults=(1, 2), seed=1234, out_scale=s)
    z, mask, cond, spk, _ = O.harness_inputs(B, T, n, seed=2, scale=s)
    dec = _decoder(64, (1, 2), p)
    zc, mc, cc, sc = z.cuda(), mask.cuda(), cond.cuda(), spk.cuda()
    torch.manual_seed(77)
    a = dec(zc, mc, cc, sc, n)
    torch.manual_seed(77)
    noise = torch.stack([torch.randn(zc.shape, device="cuda") for _ in range(n)])
    b = dec(zc, mc, cc, sc, n, noise=noise)
    assert torch.equal(a, b)


def test_host_buffer_entry_matches_device_entry():
    B, T, n, s = 2, 16, 3, 1.0 / 32
    p = O.harness_params(dim=64, dim_mults=(1, 2), seed=1234, out_scale=s)
    z, mask, cond, spk, noise = O.harness_inputs(B, T, n, seed=2, scale=s, lengths=(16, 12))
    dec = _decoder(64, (1, 2), p)
    dev = dec(z.cuda(), mask.cuda(), cond.cuda(), spk.cuda(), n, 1.0, 1.0, noise=noise.cuda())
    host = dec(z, mask, cond, spk, n, 1.0, 1.0, noise=noise)
    assert not host.is_cuda
    assert torch.equal(host, dev.cpu())


def test_argument_errors():
    from unitspeech_b200 import UnitSpeech
    p = O.harness_params(dim=64, dim_mults=(1, 2), seed=1)
    dec = _decoder(64, (1, 2), p)
    z, mask, cond, spk, noise = O.harness_inputs(1, 16, 2, seed=2)
    with pytest.raises(ValueError):
        dec(z.cuda(), mask.cuda(), cond.cuda(), spk.cuda(), 1)
    with pytest.raises(ValueError):
        dec(z[..., :15].cuda(), mask[..., :15].cuda(), cond[..., :15].cuda(), spk.cuda(), 2)
    from unitspeech_b200 import abi
    with pytest.raises(abi.UsbError):
        UnitSpeech(80, 48, (1, 2), spk_emb_dim=256).cuda()(z.cuda(), mask.cuda(), cond.cuda(), spk.cuda(), 2)


def test_state_dict_roundtrip_and_reload():
    p = O.harness_params(dim=64, dim_mults=(1, 2), seed=1234, out_scale=1 / 32)
    dec = _decoder(64, (1, 2), p)
    sd = dec.state_dict()
    assert set(sd) == set(p) and all(torch.equal(sd[k].cpu(), p[k]) for k in p)
    z, mask, cond, spk, noise = O.harness_inputs(1, 16, 2, seed=2, scale=1 / 32)
    a = dec(z.cuda(), mask.cuda(), cond.cuda(), spk.cuda(), 2, noise=noise.cuda())
    p2 = O.harness_params(dim=64, dim_mults=(1, 2), seed=99, out_scale=1 / 32)
    dec.load_state_dict(p2)
    b = dec(z.cuda(), mask.cuda(), cond.cuda(), spk.cuda(), 2, noise=noise.cuda())
    ref = O.reverse_diffusion(p2, z, mask, cond, spk, 2, noise=noise, dim=64, dim_mults=(1, 2))
    assert float((a - b).abs().max()) > 1e-4          # weights really changed
    mx, mn = _errs(b, ref)
    assert mx <= MAX_TOL and mn <= MEAN_TOL


def test_long_odd_width_utterances_and_microbatching():
    """T = 1000 (widths 1000/500/250/125: ragged tiles at every level) and automatic micro-batching of a large batch."""
    B, T, n, s = 3, 1000, 2, 1.0 / 512
    p = O.harness_params(seed=1234, out_scale=s)
    z, mask, cond, spk, noise = O.harness_inputs(B, T, n, seed=9, scale=s, lengths=(1000, 777, 1000))
    ref = O.reverse_diffusion(p, z[:2], mask[:2], cond[:2], spk[:2], n, 1.0, 1.0, noise=noise[:, :2])
    dec = _decoder(128, (1, 2, 4, 8), p)
    full = dec(z.cuda(), mask.cuda(), cond.cuda(), spk.cuda(), n, 1.0, 1.0, noise=noise.cuda())
    mx, mn = _errs(full[:2], ref)
    print(f"T=1000: rel max-abs {mx:.3e} mean-abs {mn:.3e}")
    assert mx <= MAX_TOL and mn <= MEAN_TOL
    assert float(full[1, :, 777:].abs().max()) == 0.0
    dec.max_rows_frames = 3 * 1000          # forces one utterance per library call
    chunked = dec(z.cuda(), mask.cuda(), cond.cuda(), spk.cuda(), n, 1.0, 1.0, noise=noise.cuda())
    # micro-batching is exact (bitwise batch invariance): the split-K mode is decided for the 9 x 1000 job (off), not for
    # the 3 x 1000 micro-batches (which alone would run in latency mode)
    assert torch.equal(chunked, full)
    alone = dec(z[:1].cuda(), mask[:1].cuda(), cond[:1].cuda(), spk[:1].cuda(), n, 1.0, 1.0, noise=noise[:, :1].cuda())
    mx1, _ = _errs(alone, ref[:1])
    assert mx1 <= MAX_TOL and float((alone - full[:1]).abs().max()) <= 2e-3   # latency mode: same result to fp32 rounding


def test_fused_mel_denormalisation_is_bit_exact():
    """(y + 1) / 2 * (mel_max - mel_min) + mel_min (inference.py:140) fused into the last sampler step."""
    from unitspeech_b200 import denormalize_mel
    B, T, n, s = 2, 16, 3, 1.0 / 32
    p = O.harness_params(dim=64, dim_mults=(1, 2), seed=1234, out_scale=s)
    z, mask, cond, spk, noise = O.harness_inputs(B, T, n, seed=2, scale=s, lengths=(16, 12))
    dec = _decoder(64, (1, 2), p)
    g = torch.Generator().manual_seed(3)
    mel_min = -11.0 + torch.rand(80, 1, generator=g)
    mel_max = 1.5 + torch.rand(80, 1, generator=g)
    args = (z.cuda(), mask.cuda(), cond.cuda(), spk.cuda(), n, 1.0, 1.0)
    y = dec(*args, noise=noise.cuda())
    fused = dec(*args, noise=noise.cuda(), denorm=(mel_min, mel_max))
    assert torch.equal(fused, denormalize_mel(y, mel_min.cuda(), mel_max.cuda()))
    # (1, 80, 1)-shaped ranges as scripts/finetune.py:106-107 stores them, host entry, and switching it off again
    fused_h = dec(z, mask, cond, spk, n, 1.0, 1.0, noise=noise, denorm=(mel_min.view(1, 80, 1), mel_max.view(1, 80, 1)))
    assert torch.equal(fused_h, fused.cpu())
    assert torch.equal(dec(*args, noise=noise.cuda()), y)
    with pytest.raises(ValueError):
        dec(*args, noise=noise.cuda(), denorm=(mel_min[:10], mel_max[:10]))


def test_fp16_saturation_is_clamped_and_reported():
    """out-of-range activations saturate at +-65504 (never inf) and the clamp is counted."""
    B, T, n, s = 1, 16, 2, 1.0 / 32
    p = O.harness_params(dim=64, dim_mults=(1, 2), seed=1234, out_scale=s)
    z, mask, cond, spk, noise = O.harness_inputs(B, T, n, seed=2, scale=s)
    dec = _decoder(64, (1, 2), p)
    out = dec(z.cuda(), mask.cuda(), cond.cuda(), spk.cuda(), n, 1.0, 1.0, noise=noise.cuda())
    assert dec.saturation_count() == 0 and torch.isfinite(out).all()
    # inputs scaled by 1e6 drive the first conv's fp16 outputs past the format's range
    dec(z.cuda() * 1e6, mask.cuda(), cond.cuda() * 1e6, spk.cuda(), n, 1.0, 1.0, noise=noise.cuda())
    assert dec.saturation_count() > 0
    assert dec.saturation_count(reset=True) > 0 and dec.saturation_count() == 0


class _StubEncoder(torch.nn.Module):
    """Deterministic stand-in for the text/unit encoder (unitspeech/encoder.py:294): (cond_x, x, x_mask)."""

    def __init__(self, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.table = torch.nn.Parameter(torch.randn(50, 80, generator=g).clamp(-1, 1))

    def forward(self, tokens, lengths):
        cond_x = self.table[tokens].transpose(1, 2)                                   # (B, 80, L)
        L = tokens.shape[1]
        x_mask = (torch.arange(L, device=tokens.device)[None] < lengths[:, None]).unsqueeze(1).to(cond_x.dtype)
        return cond_x * x_mask, cond_x * x_mask, x_mask


class _StubDuration(torch.nn.Module):
    def forward(self, x, x_mask, w=None, g=None, reverse=True):
        # log-durations 3..6 frames per token, depends on the token content
        return torch.log(3.0 + 3.0 * torch.sigmoid(x[:, :1] * 4)) * x_mask


def test_execute_text_to_speech_matches_reference_glue():
    """execute_text_to_speech (unitspeech.py:414-450): durations -> alignment -> cond_y -> z -> sampler -> crop."""
    from unitspeech_b200 import fix_len_compatibility, generate_path, sequence_mask
    s = 1.0 / 32
    p = O.harness_params(dim=64, dim_mults=(1, 2), seed=1234, out_scale=s)
    dec = _decoder(64, (1, 2), p)
    enc, dur = _StubEncoder().cuda(), _StubDuration().cuda()
    tokens = torch.tensor([[3, 7, 11, 2, 40, 5]]).cuda()
    lengths = torch.tensor([6]).cuda()
    spk = torch.nn.functional.normalize(torch.randn(1, 1, 256, generator=torch.Generator().manual_seed(4)), dim=-1).cuda()
    n = 3
    torch.manual_seed(5)
    y_enc, y_dec, attn = dec.execute_text_to_speech(tokens, lengths, spk, enc, dur, num_downsamplings_in_unet=1,
                                                    diffusion_steps=n, length_scale=1.0, text_gradient_scale=1.0,
                                                    spk_gradient_scale=1.0)
    # the same glue restated with the oracle sampler
    with torch.no_grad():
        cond_x, x, x_mask = enc(tokens, lengths)
        w_ceil = torch.ceil(torch.exp(dur(x, x_mask)) * x_mask)
        y_len = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
        T = fix_len_compatibility(int(y_len.max()), 1)
        y_mask = sequence_mask(y_len, T).unsqueeze(1).to(x_mask.dtype)
        path = generate_path(w_ceil.squeeze(1), (x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)).squeeze(1))
        cond_y = torch.matmul(path.transpose(1, 2), cond_x.transpose(1, 2)).transpose(1, 2).contiguous()
    torch.manual_seed(5)
    z = torch.randn_like(cond_y)
    noise = torch.stack([torch.randn(z.shape, device="cuda") for _ in range(n)])
    ref = O.reverse_diffusion(p, z.cpu(), y_mask.cpu(), cond_y.cpu(), spk.cpu(), n, 1.0, 1.0, noise=noise.cpu(),
                              dim=64, dim_mults=(1, 2))[:, :, :int(y_len.max())]
    # (the reference slices attn[:, :, :y_max] on the TOKEN axis of the (B,1,Tx,Ty) map, unitspeech.py:450 -- kept)
    assert y_dec.shape == ref.shape and tuple(attn.shape) == (1, 1, 6, T)
    assert torch.equal(y_enc, cond_y[:, :, :int(y_len.max())])
    mx, mn = _errs(y_dec, ref)
    assert mx <= MAX_TOL and mn <= MEAN_TOL


def test_on_device_front_end_matches_reference_glue():
    """execute_text_to_speech(max_frames=...): the durations -> mask / path / cond_y glue in one
    kernel at a fixed frame capacity, against the reference's torch glue (unitspeech.py:421-441) on a ragged batch."""
    from unitspeech_b200 import fix_len_compatibility, generate_path, sequence_mask
    p = O.harness_params(dim=64, dim_mults=(1, 2), seed=1234, out_scale=1.0 / 32)
    dec = _decoder(64, (1, 2), p)
    enc, dur = _StubEncoder().cuda(), _StubDuration().cuda()
    tokens = torch.tensor([[3, 7, 11, 2, 40, 5, 9], [8, 1, 30, 4, 0, 0, 0]]).cuda()
    lengths = torch.tensor([7, 4]).cuda()
    spk = torch.nn.functional.normalize(torch.randn(2, 1, 256, generator=torch.Generator().manual_seed(4)), dim=-1).cuda()
    n = 3
    with torch.no_grad():
        cond_x, x, x_mask = enc(tokens, lengths)
        w_ceil = torch.ceil(torch.exp(dur(x, x_mask)) * x_mask)
        y_len = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
        T = fix_len_compatibility(int(y_len.max()), 1)
        y_mask = sequence_mask(y_len, T).unsqueeze(1).to(x_mask.dtype)
        path = generate_path(w_ceil.squeeze(1), (x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)).squeeze(1))
        cond_y = torch.matmul(path.transpose(1, 2), cond_x.transpose(1, 2)).transpose(1, 2).contiguous()
    noise = torch.randn(n, 2, 80, T, generator=torch.Generator().manual_seed(9)).cuda() / 32
    torch.manual_seed(5)
    y_enc, y_dec, attn = dec.execute_text_to_speech(tokens, lengths, spk, enc, dur, num_downsamplings_in_unet=1,
                                                    diffusion_steps=n, text_gradient_scale=1.0, spk_gradient_scale=1.0,
                                                    noise=noise, max_frames=T)
    assert torch.equal(dec.last_y_lengths, y_len)
    assert torch.equal(y_enc, cond_y) and torch.equal(attn.squeeze(1), path)
    torch.manual_seed(5)
    z = torch.randn_like(cond_y)
    want = dec.forward(z, y_mask, cond_y, spk, n_timesteps=n, text_gradient_scale=1.0, spk_gradient_scale=1.0, noise=noise)
    assert torch.equal(y_dec, want)
    # a larger capacity only appends masked (zero) frames
    torch.manual_seed(5)
    y_enc2, y_dec2, attn2 = dec.execute_text_to_speech(tokens, lengths, spk, enc, dur, num_downsamplings_in_unet=1,
                                                       diffusion_steps=n, max_frames=T + 6)
    assert y_dec2.shape[-1] == T + 6 and float(y_dec2[:, :, T:].abs().max()) == 0.0 and float(y_enc2[:, :, T:].abs().max()) == 0.0
    assert torch.equal(attn2[..., :T].squeeze(1), path) and float(attn2[..., T:].abs().max()) == 0.0


# ---------------------------------------------------------------------------------------------------------------------
# fine-tuning objective, forward value (forward half)
# ---------------------------------------------------------------------------------------------------------------------
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


@pytest.mark.parametrize("name", ["loss_d64", "loss_full"])
def test_loss_t_matches_reference_golden_and_oracle(golden_dir, name):
    from make_golden_loss import LOSS_CASES, loss_inputs
    dim, mults, B, T, lengths, ts, s = LOSS_CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    params = O.harness_params(dim=dim, dim_mults=mults, seed=1234, out_scale=s)
    x0, mask, cond, spk = loss_inputs(B, T, lengths)
    t = torch.tensor(ts)
    dec = _decoder(dim, mults, params)
    torch.manual_seed(77)                       # same global-RNG draw as the reference's forward_diffusion (CPU tensors in)
    with torch.no_grad():                       # forward value through the fused usb_loss_t entry
        loss, xt = dec.loss_t(x0, mask, cond, t, spk)
    assert not loss.requires_grad and loss.dim() == 0
    assert float((xt - torch.from_numpy(g["xt"])).abs().max()) <= 1e-5
    rel = abs(float(loss) - float(g["loss"])) / float(g["loss"])
    print(f"{name}: loss {float(loss):.6f} reference {float(g['loss']):.6f} rel {rel:.2e}")
    assert rel <= 2e-3
    # forward_diffusion alone, same draw
    torch.manual_seed(77)
    xt2, zm = dec.forward_diffusion(x0, mask, t)
    torch.manual_seed(77)
    z = torch.randn(x0.shape)
    assert torch.equal(xt2, xt) and torch.equal(zm, z * mask)
    # device tensors in -> device tensors out, and repeatable given the generator state
    with torch.no_grad():
        torch.cuda.manual_seed(5)
        l1, _ = dec.loss_t(x0.cuda(), mask.cuda(), cond.cuda(), t.cuda(), spk.cuda())
        torch.cuda.manual_seed(5)
        l2, _ = dec.loss_t(x0.cuda(), mask.cuda(), cond.cuda(), t.cuda(), spk.cuda())
    assert l1.is_cuda and torch.equal(l1, l2)
    # with autograd enabled the same call goes through the fine-tune engine and carries a graph
    torch.manual_seed(77)
    l3, xt3 = dec.loss_t(x0, mask, cond, t, spk)
    assert l3.requires_grad and abs(float(l3) - float(loss)) <= 1e-3 * float(loss)
    assert float((xt3 - xt).abs().max()) <= 1e-6


@pytest.mark.parametrize("name", ["finetune_d64", "finetune_d64_short"])
def test_fine_tune_objective_matches_reference_golden(golden_dir, name):
    import random
    from make_golden_loss import FT_CASES, FT_OUT_SCALE, finetune_inputs
    dim, mults, B, Lt, T, y_lengths, seg = FT_CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    params = O.harness_params(dim=dim, dim_mults=mults, seed=1234, out_scale=FT_OUT_SCALE)
    cond_x, y, y_mask, yl, attn, spk = finetune_inputs(B, Lt, T, y_lengths)
    dec = _decoder(dim, mults, params)
    random.seed(5)
    torch.manual_seed(78)
    loss = dec.fine_tune(cond_x, y, y_mask, yl, T, attn, spk, seg, 80)
    rel = abs(float(loss) - float(g["loss"])) / float(g["loss"])
    print(f"{name}: loss {float(loss):.6f} reference {float(g['loss']):.6f} rel {rel:.2e}")
    assert rel <= 2e-3
