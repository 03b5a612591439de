"""B200 parity tests of the vocoder stage (usb_vocoder_* through unitspeech_b200.BigVGAN) against the oracle
(oracle/bigvgan_oracle.py) and the golden vectors produced by the unmodified reference generator.

Tolerance: the waveform lives in [-1, 1]; activations are stored in fp16 between kernels, accumulation is fp32.
The bar written here is max-abs <= 1e-2 and mean-abs <= 1e-3 against the fp32 reference (the mel bar of the
decoder, applied to the audio)."""

import ctypes
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from oracle import bigvgan_oracle as V  # noqa: E402

from make_golden_vocoder import CONFIGS  # noqa: E402

pytestmark = pytest.mark.gpu

MAX_ABS, MEAN_ABS = 1e-2, 1e-3


def _mel(h, B, T, seed=17):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, h["num_mels"], T, generator=g) * 2 - 4


def _vocoder(h, seed=4321):
    from unitspeech_b200 import BigVGAN
    voc = BigVGAN(h)
    voc.load_state_dict(V.harness_params(h, seed))
    return voc.cuda().eval()


def _report(name, out, ref):
    d = (out - ref).abs()
    print(f"{name}: max-abs {float(d.max()):.3e} mean-abs {float(d.mean()):.3e} ref-absmax {float(ref.abs().max()):.3f}")
    return float(d.max()), float(d.mean())


@pytest.mark.parametrize("N,L,C,Cr", [(1, 1, 64, 64), (2, 5, 64, 24), (1, 64, 128, 96), (2, 200, 64, 48), (1, 777, 192, 192),
                                       (1, 130, 64, 8)])
def test_snake_activation_kernel(N, L, C, Cr):
    from unitspeech_b200 import abi
    lib = abi.load_library()
    g = torch.Generator().manual_seed(L * 7 + C)
    x = (torch.randn(N, C, L, generator=g) * 1.5).half().float()
    alpha, beta = torch.randn(C, generator=g) * 0.4, torch.randn(C, generator=g) * 0.4
    p = {"a.act.alpha": alpha, "a.act.beta": beta}
    filt = V.kaiser_sinc_filter1d(0.25, 0.3, 12)
    ref = V.activation1d(p, "a", x, filt, dict(activation="snakebeta", snake_logscale=True))
    xd = x.permute(0, 2, 1).contiguous().cuda().half()
    xd[:, :, Cr:] = float("nan")             # padding channels may hold anything: they must not be read
    out = torch.full_like(xd, float("nan"))
    al, ib = torch.exp(alpha).cuda(), (1.0 / (torch.exp(beta) + 1e-9)).cuda()
    abi.check(lib.usb_op_snake_act(xd.data_ptr(), al.data_ptr(), ib.data_ptr(), N, L, C, Cr, out.data_ptr(),
                                   int(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    got = out.float().cpu().permute(0, 2, 1)
    assert torch.isfinite(got).all()
    assert float(got[:, Cr:].abs().max()) == 0.0 if Cr < C else True
    got, ref = got[:, :Cr], ref[:, :Cr]
    err = (got - ref).abs()
    assert float(err.max()) <= 4e-3 * max(1.0, float(ref.abs().max())), float(err.max())


# (padded c_in, real c_in, padded c_out, real c_out, k, dilation, N, L, residual)
CONV1D_CASES = [
    (64, 24, 64, 24, 3, 1, 2, 300, False),       # resident weights, 2 of 4 K slices, ragged last tile
    (64, 24, 64, 24, 11, 5, 1, 1000, True),      # widest halo (50 positions), in-place residual
    (64, 48, 64, 48, 7, 3, 2, 129, True),        # 3 of 4 K slices, one position past a tile
    (64, 64, 64, 64, 11, 1, 1, 17, False),       # signal shorter than the halo
    (128, 96, 128, 96, 3, 5, 1, 515, True),      # two chunks, the second one half padding, resident weights
    (128, 96, 128, 96, 11, 3, 2, 260, True),     # streamed weights, BN = 128
    (192, 192, 192, 192, 7, 5, 1, 400, False),   # BN = 192 (three epilogue slabs over two groups)
    (192, 192, 192, 192, 11, 1, 2, 131, True),
    (384, 384, 384, 384, 3, 3, 1, 257, True),    # two N tiles of 192
    (256, 256, 256, 256, 7, 1, 1, 384, True),    # BN = 256, two activation buffers
    (128, 128, 128, 128, 7, 3, 1, 200, False),   # no residual, Cout % 128 == 0: the swapped-operand kernel
    (64, 64, 64, 64, 1, 1, 1, 100, True),        # k = 1: the per-tap kernel
]


@pytest.mark.parametrize("ci,cir,co,cor,k,d,N,L,with_res", CONV1D_CASES)
def test_conv1d_layer_matches_torch(ci, cir, co, cor, k, d, N, L, with_res):
    """One generator Conv1d (vocoder/models.py:46-58, "same" padding of xutils.get_padding) through usb_op_conv1d against
    torch.nn.functional.conv1d in fp32 on the same fp16-rounded operands: every kernel the layer can pick (1-D halo kernel
    with resident / streamed weights and N = 64 / 128 / 192 / 256, swapped-operand kernel, per-tap kernel), ragged lengths,
    channel padding, in-place residual."""
    from unitspeech_b200 import abi
    lib = abi.load_library()
    g = torch.Generator().manual_seed(ci * 131 + k * 17 + d + L)
    x = torch.zeros(N, ci, L)
    x[:, :cir] = torch.randn(N, cir, L, generator=g)
    x = x.half().float()
    w = torch.zeros(co, ci, k)
    w[:cor, :cir] = torch.randn(cor, cir, k, generator=g) / (cir * k) ** 0.5
    w = w.half().float()
    bias = torch.zeros(co)
    bias[:cor] = torch.randn(cor, generator=g) * 0.1
    res = torch.zeros(N, co, L)
    res[:, :cor] = torch.randn(N, cor, L, generator=g)
    res = res.half().float()
    ref = torch.nn.functional.conv1d(x, w, bias, padding=(k * d - d) // 2, dilation=d)
    if with_res:
        ref = ref + res
    xd = x.permute(0, 2, 1).contiguous().cuda().half()
    wd = w.permute(0, 2, 1).contiguous().reshape(co, k * ci).cuda().half()      # [co][t * ci + c]
    bd = bias.cuda()
    out = res.permute(0, 2, 1).contiguous().cuda().half() if with_res else torch.full((N, L, co), float("nan"), device="cuda",
                                                                                    dtype=torch.half)
    abi.check(lib.usb_op_conv1d(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr() if with_res else None, out.data_ptr(),
                                N, L, ci, cir, co, k, d, int(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    got = out.float().cpu().permute(0, 2, 1)
    assert torch.isfinite(got).all()
    assert float(got[:, cor:].abs().max()) == 0.0 if cor < co else True      # padding channels stay exactly zero
    err = (got - ref).abs()
    tol = 2e-3 * max(1.0, float(ref.abs().max()))      # fp16 output rounding (2^-11 relative) + fp32 accumulation order
    assert float(err.max()) <= tol, (float(err.max()), tol)


def test_library_filter_equals_reference_filter(golden_dir):
    from unitspeech_b200 import abi
    buf = (ctypes.c_float * 12)()
    abi.check(abi.load_library().usb_vocoder_filter(buf))
    g = np.load(os.path.join(golden_dir, "vocoder_small.npz"))
    assert np.abs(np.array(buf[:], dtype=np.float32) - g["filt"]).max() <= 1e-7


@pytest.mark.parametrize("name", list(CONFIGS))
def test_vocoder_matches_reference_golden(golden_dir, name):
    h, B, T = CONFIGS[name]
    ref = torch.from_numpy(np.load(os.path.join(golden_dir, name + ".npz"))["out"])
    voc = _vocoder(h)
    out = voc(_mel(h, B, T).cuda()).cpu()
    assert out.shape == ref.shape
    mx, mn = _report(name, out, ref)
    assert mx <= MAX_ABS and mn <= MEAN_ABS


def test_vocoder_public_config_batch_vs_oracle():
    h = dict(V.PUBLIC_22KHZ_80BAND)
    B, T = 2, 40
    mel = _mel(h, B, T, seed=5)
    ref = V.bigvgan_forward(V.harness_params(h), mel, h)
    voc = _vocoder(h)
    out = voc(mel.cuda()).cpu()
    assert out.shape == (B, 1, T * 256)
    mx, mn = _report("public B2 T40", out, ref)
    assert mx <= MAX_ABS and mn <= MEAN_ABS
    # an utterance is synthesised identically alone and inside a batch; a second call repeats bit for bit
    solo = voc(mel[1:2].cuda()).cpu()
    assert torch.equal(solo[0], out[1])
    assert torch.equal(voc(mel.cuda()).cpu(), out)
    assert voc.launch_count > 0 and voc.workspace_bytes > 0 and voc.flops_per_call > 0


def test_vocoder_host_entry_microbatching_and_weight_norm_checkpoint():
    h = dict(V.PUBLIC_22KHZ_80BAND, upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4], upsample_initial_channel=128)
    mel = _mel(h, 3, 37, seed=9)
    voc = _vocoder(h)
    dev_out = voc(mel.cuda()).cpu()
    host_out = voc(mel)                      # CPU tensor in -> usb_vocoder_forward_host -> CPU tensor out
    assert not host_out.is_cuda and torch.equal(host_out, dev_out)
    voc.max_frames_per_call = 40             # forces one utterance per call
    assert torch.equal(voc(mel.cuda()).cpu(), dev_out)
    # a reference-format checkpoint (weight_g / weight_v + filter buffers) loads to the same generator
    from unitspeech_b200 import BigVGAN
    p = V.harness_params(h)
    ck = {}
    for k, v in p.items():
        if k.endswith(".weight"):
            g = v.flatten(1).norm(dim=1).view(-1, 1, 1)
            ck[k + "_g"], ck[k + "_v"] = g, v * 3.0          # any positive rescaling of v folds back to v * g/||v||
        else:
            ck[k] = v
    ck["activation_post.upsample.filter"] = torch.zeros(1, 1, 12)
    voc2 = BigVGAN(h)
    voc2.load_state_dict(ck)
    voc2.cuda().eval().remove_weight_norm()
    d = (voc2(mel.cuda()).cpu() - dev_out).abs().max()
    assert float(d) <= 2e-3


def test_vocoder_errors_are_loud():
    from unitspeech_b200 import BigVGAN, abi
    h = dict(V.PUBLIC_22KHZ_80BAND, upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4], upsample_initial_channel=128)
    voc = _vocoder(h)
    with pytest.raises(ValueError):
        voc(torch.zeros(1, 79, 8).cuda())
    bad = dict(h, upsample_kernel_sizes=[7, 4])
    v2 = BigVGAN(bad).cuda()
    with pytest.raises(abi.UsbError):
        v2(torch.zeros(1, 80, 8).cuda())


def test_decoder_then_vocoder_pipeline_matches_oracle_pipeline():
    """inference.py:128-141: decoder -> de-normalise -> vocoder -> clamp, both stages on the CUDA path vs both oracles."""
    from oracle import unitspeech_oracle as O
    from unitspeech_b200 import UnitSpeech, denormalize_mel
    params = O.harness_params(dim=64, dim_mults=(1, 2), seed=1234)
    B, T, n = 2, 16, 4
    z, mask, cond, spk, noise = O.harness_inputs(B, T, n, seed=3, scale=1.0 / 512, lengths=[16, 12])
    y_ref = torch.cat([O.reverse_diffusion(params, z[b:b + 1], mask[b:b + 1], cond[b:b + 1], spk[b:b + 1], n, 1.0, 1.0,
                                           noise=noise[:, b:b + 1], dim=64, dim_mults=(1, 2)) for b in range(B)])
    h = dict(V.PUBLIC_22KHZ_80BAND, upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4], upsample_initial_channel=128)
    vp = V.harness_params(h)
    mel_min, mel_max = torch.full((80, 1), -11.5), torch.full((80, 1), 2.0)
    wav_ref = V.bigvgan_forward(vp, (y_ref + 1) / 2 * (mel_max - mel_min) + mel_min, h).squeeze(1).clamp(-1, 1)

    dec = UnitSpeech(80, 64, (1, 2), spk_emb_dim=256)
    dec.load_state_dict(params)
    dec = dec.cuda().eval()
    voc = _vocoder(h)
    y = dec(z.cuda(), mask.cuda(), cond.cuda(), spk.cuda(), n, text_gradient_scale=1.0, spk_gradient_scale=1.0,
            noise=noise.cuda())
    wav = voc(denormalize_mel(y, mel_min.cuda(), mel_max.cuda())).squeeze(1).clamp(-1, 1).cpu()
    assert wav.shape == (B, T * 8)
    mx, mn = _report("pipeline", wav, wav_ref)
    assert mx <= MAX_ABS and mn <= MEAN_ABS
    # the same pipeline with the de-normalisation fused on either side: into the vocoder's input pack ...
    wav_v = voc(y, mel_min=mel_min, mel_max=mel_max).squeeze(1).clamp(-1, 1).cpu()
    assert torch.equal(wav_v, wav)
    # ... or into the sampler's last step
    y_dn = dec(z.cuda(), mask.cuda(), cond.cuda(), spk.cuda(), n, text_gradient_scale=1.0, spk_gradient_scale=1.0,
               noise=noise.cuda(), denorm=(mel_min, mel_max))
    assert torch.equal(voc(y_dn).squeeze(1).clamp(-1, 1).cpu(), wav)
