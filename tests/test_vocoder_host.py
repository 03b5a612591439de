"""CPU-side checks of the vocoder drop-in: state_dict surface, weight-norm folding, ABI struct, loud failure."""

import ctypes

import pytest
import torch

from oracle import bigvgan_oracle as V


def test_vocoder_state_dict_surface_matches_reference_generator():
    from unitspeech_b200 import BigVGAN
    for h in (V.PUBLIC_22KHZ_80BAND,
              dict(V.PUBLIC_22KHZ_80BAND, resblock="2", activation="snake", snake_logscale=False,
                   upsample_initial_channel=256)):
        voc = BigVGAN(dict(h))
        got = {k: tuple(v.shape) for k, v in voc.state_dict().items()}
        assert got == V.param_shapes(h)            # the keys/shapes golden-checked against the reference class
    assert sum(v.numel() for v in BigVGAN(dict(V.PUBLIC_22KHZ_80BAND)).state_dict().values()) == 112199473


def test_weight_norm_folding_equals_torch_remove_weight_norm():
    from unitspeech_b200.vocoder import fold_weight_norm
    torch.manual_seed(0)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        conv = torch.nn.utils.weight_norm(torch.nn.Conv1d(6, 4, 3))
        convt = torch.nn.utils.weight_norm(torch.nn.ConvTranspose1d(6, 4, 4, 2))
    for m in (conv, convt):
        with torch.no_grad():
            m.weight_g.mul_(torch.rand_like(m.weight_g) + 0.5)
    sd = {"a." + k: v.detach().clone() for k, v in conv.state_dict().items()}
    sd.update({"b." + k: v.detach().clone() for k, v in convt.state_dict().items()})
    sd["a.up.filter"] = torch.zeros(1, 1, 12)
    folded = fold_weight_norm(sd)
    torch.nn.utils.remove_weight_norm(conv)
    torch.nn.utils.remove_weight_norm(convt)
    assert set(folded) == {"a.weight", "a.bias", "b.weight", "b.bias"}
    assert torch.allclose(folded["a.weight"], conv.weight, atol=1e-6)
    assert torch.allclose(folded["b.weight"], convt.weight, atol=1e-6)


def test_synthetic_vocoder_state_equals_parity_harness():
    from unitspeech_b200.synthetic import PUBLIC_VOCODER_CONFIG, vocoder_state
    assert PUBLIC_VOCODER_CONFIG == V.PUBLIC_22KHZ_80BAND
    h = dict(V.PUBLIC_22KHZ_80BAND, upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4], upsample_initial_channel=128)
    a, b = vocoder_state(h), V.harness_params(h)
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)


def test_vocoder_config_struct_and_loud_failure_without_gpu():
    from unitspeech_b200 import BigVGAN, abi, build
    build.build_library()
    # 2 + 8 + 8 + 3 + 4 + 1 + 16 + 3 int32, no padding
    assert ctypes.sizeof(abi.UsbVocoderConfig) == 45 * 4
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    voc = BigVGAN(dict(V.PUBLIC_22KHZ_80BAND, upsample_rates=[2], upsample_kernel_sizes=[4], upsample_initial_channel=64))
    with pytest.raises(abi.UsbError):
        voc(torch.zeros(1, 80, 4))
    cfg = abi.UsbVocoderConfig()
    cfg.num_mels, cfg.n_upsamples, cfg.upsample_initial_channel = 80, 1, 64
    cfg.resblock_type, cfg.n_resblock_kernels, cfg.n_dilations, cfg.activation = 1, 1, 1, 1
    hp = ctypes.c_void_p()
    assert abi.load_library().usb_vocoder_create(ctypes.byref(cfg), ctypes.byref(hp)) != 0 and not hp.value
