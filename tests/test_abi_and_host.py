"""CPU-side checks: the C-ABI library builds, loads and exports exactly what include/unitspeech_b200.h declares;
the host-side mirror of the reference interface (schedule, helpers, state_dict surface) behaves like the reference."""

import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import unitspeech_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from unitspeech_b200 import abi, build
    build.build_library()          # nvcc cross-compiles without a GPU; no-op when up to date
    return abi.load_library()


def _header_functions():
    text = "".join(open(os.path.join(ROOT, "include", f)).read() for f in sorted(os.listdir(os.path.join(ROOT, "include")))
                   if f.endswith(".h"))
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(usb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    from unitspeech_b200 import abi
    names = _header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in abi.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(abi.SIGNATURES) == names, "ctypes table and header disagree"
    assert lib.usb_version() >= 1


def test_config_struct_layout_matches_header():
    from unitspeech_b200 import abi
    # 3 int32 + int32[8] + 2 int32 + 3 float + int32 = 17 * 4 bytes, no padding
    assert ctypes.sizeof(abi.UsbConfig) == 17 * 4
    assert abi.UsbConfig.dim_mults.offset == 12 and abi.UsbConfig.device.offset == 64


def test_create_fails_loudly_without_a_b200(lib):
    from unitspeech_b200 import abi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = abi.UsbConfig()
    cfg.n_feats, cfg.dim, cfg.n_mults, cfg.groups, cfg.spk_emb_dim, cfg.device = 80, 128, 4, 8, 256, 0
    for i, m in enumerate((1, 2, 4, 8)):
        cfg.dim_mults[i] = m
    h = ctypes.c_void_p()
    rc = lib.usb_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc != 0 and not h.value
    assert len(lib.usb_last_error()) > 0
    with pytest.raises(abi.UsbError):
        abi.check(rc)


def test_decoder_has_no_cpu_fallback():
    from unitspeech_b200 import UnitSpeech, abi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    dec = UnitSpeech(80, 64, (1, 2), spk_emb_dim=256)
    z, mask, cond, spk, noise = O.harness_inputs(1, 16, 2)
    with pytest.raises(abi.UsbError):
        dec(z, mask, cond, spk, 2, noise=noise)


def test_state_dict_surface_matches_reference():
    from unitspeech_b200 import UnitSpeech
    dec = UnitSpeech(n_feats=80, dim=128, dim_mults=(1, 2, 4, 8), beta_min=0.05, beta_max=20, pe_scale=1000,
                     spk_emb_dim=256)
    shapes = O.param_shapes(80, 128, (1, 2, 4, 8), 256)      # pinned to the reference by tests/golden (strict load)
    sd = dec.state_dict()
    assert len(sd) == 230 and set(sd) == set(shapes)
    assert all(tuple(sd[k].shape) == shapes[k] for k in sd)
    assert dec.nparams == 119145177
    # reference zero-inits: Rezero g, text_uncon, spk_uncon (unitspeech.py:40,230-231)
    assert float(sd["estimator.mid_attn.fn.g"].abs().sum()) == 0 and float(sd["spk_uncon"].abs().sum()) == 0
    p = O.harness_params()
    dec.load_state_dict(p, strict=True)
    assert all(torch.equal(dec.state_dict()[k], p[k]) for k in p)
    if not torch.cuda.is_available():      # the training objective runs on the CUDA library too: no CPU fallback
        from unitspeech_b200 import abi
        with pytest.raises(abi.UsbError):
            dec.compute_loss(torch.zeros(1, 80, 8), torch.ones(1, 1, 8), torch.zeros(1, 80, 8), torch.zeros(1, 1, 256))


@pytest.mark.parametrize("n", [2, 4, 50, 500])
def test_product_schedule_equals_oracle(n):
    from unitspeech_b200 import schedule
    assert torch.equal(schedule.step_coefficients(n, 0.05, 20.0), O.step_coefficients(n, 0.05, 20.0))
    assert torch.equal(schedule.step_times(n), O.step_times(n))
    with pytest.raises(ValueError):
        schedule.schedule_tables(1, 0.05, 20.0)


def test_posemb_freqs_match_reference_table(golden_dir):
    from unitspeech_b200 import schedule
    g = np.load(os.path.join(golden_dir, "schedule.npz"))
    t = torch.from_numpy(g["posemb_t"])
    f = schedule.posemb_freqs(128)
    arg = 1000 * t.unsqueeze(1) * f.unsqueeze(0)
    emb = torch.cat((arg.sin(), arg.cos()), -1)
    assert np.array_equal(emb.numpy(), g["posemb_128"])


def test_helpers_follow_reference_semantics():
    from unitspeech_b200 import fix_len_compatibility, generate_path, sequence_mask
    assert [fix_len_compatibility(v) for v in (1, 8, 9, 255, 256)] == [8, 8, 16, 256, 256]
    assert fix_len_compatibility(5, 2) == 8
    m = sequence_mask(torch.tensor([3, 1, 4]))
    assert m.tolist() == [[True, True, True, False], [True, False, False, False], [True, True, True, True]]
    dur = torch.tensor([[2.0, 0.0, 3.0]])
    mask = torch.ones(1, 3, 6)
    path = generate_path(dur, mask)
    assert path[0].tolist() == [[1, 1, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0], [0, 0, 1, 1, 1, 0]]


def test_mel_normalisation_contract_roundtrip():
    from unitspeech_b200 import denormalize_mel, normalize_mel
    g = torch.Generator().manual_seed(0)
    mel_min, mel_max = -torch.rand(80, 1, generator=g) * 10 - 2, torch.rand(80, 1, generator=g) * 2
    y = torch.rand(2, 80, 16, generator=g) * 2 - 1
    mel = denormalize_mel(y, mel_min, mel_max)                      # inference.py:140
    assert torch.allclose(mel, (y + 1) / 2 * (mel_max - mel_min) + mel_min)
    assert torch.allclose(normalize_mel(mel, mel_min, mel_max), y, atol=1e-5)


def test_checkpoint_roundtrip_in_reference_format(tmp_path):
    """{"model", "spk_emb", "mel_min", "mel_max", "iteration"} (train_STEP1.py:297-304) written, read back, arch inferred."""
    from unitspeech_b200 import UnitSpeech, load_decoder_checkpoint, save_decoder_checkpoint
    p = O.harness_params(dim=64, dim_mults=(1, 2, 4), seed=3)
    dec = UnitSpeech(80, 64, (1, 2, 4), spk_emb_dim=256)
    dec.load_state_dict(p)
    g = torch.Generator().manual_seed(0)
    spk, mn, mx = torch.randn(1, 256, generator=g), -torch.rand(80, 1, generator=g) * 8, torch.rand(80, 1, generator=g)
    f = tmp_path / "dec.pt"
    save_decoder_checkpoint(f, dec, spk_emb=spk, mel_min=mn, mel_max=mx, iteration=123)
    raw = torch.load(f)
    assert set(raw) == {"model", "spk_emb", "mel_min", "mel_max", "iteration"} and set(raw["model"]) == set(p)
    b = load_decoder_checkpoint(f, device="cpu")
    assert (b.decoder.n_feats, b.decoder.dim, b.decoder.dim_mults, b.decoder.spk_emb_dim) == (80, 64, (1, 2, 4), 256)
    assert all(torch.equal(b.decoder.state_dict()[k], p[k]) for k in p)
    assert b.iteration == 123 and torch.equal(b.spk_emb, spk)
    y = torch.rand(1, 80, 8, generator=g) * 2 - 1
    assert torch.allclose(b.denormalize(y), (y + 1) / 2 * (mx - mn) + mn)


def test_loss_t_refuses_gradients_it_cannot_produce():
    """the CUDA backward pass yields decoder-parameter gradients only; a caller that expects the diffusion
    loss to reach its encoder through cond / x0 / spk_emb (train_STEP1.py:381, train_STEP2.py:299) must get an error, not a
    silently constant loss."""
    from unitspeech_b200 import UnitSpeech
    dec = UnitSpeech(80, 64, (1, 2), spk_emb_dim=256)
    x0, cond = torch.zeros(1, 80, 16), torch.zeros(1, 80, 16, requires_grad=True)
    mask, t, spk = torch.ones(1, 1, 16), torch.tensor([0.5]), torch.zeros(1, 1, 256)
    with pytest.raises(NotImplementedError, match="decoder parameters only"):
        dec.loss_t(x0, mask, cond, t, spk)
    for p in dec.parameters():
        p.requires_grad_(False)                      # frozen decoder (train_STEP2.py) makes no difference
    with pytest.raises(NotImplementedError):
        dec.compute_loss(x0, mask, cond, spk_emb=spk)


def test_installed_reference_is_the_unmodified_reference():
    """baseline/_ref (scripts/install_ref.py) holds byte-identical copies of the reference's decoder-path files."""
    import filecmp
    import sys
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import install_ref
    from oracle import ref_shim
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference tree not present (GPU box)")
    assert install_ref.install(verbose=False)
    assert ref_shim.installed_reference_available()
    for f in install_ref.FILES:
        assert filecmp.cmp(os.path.join("/root/reference", f), os.path.join(install_ref.DEST, f), shallow=False), f
