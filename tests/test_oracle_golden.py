"""Pins oracle/unitspeech_oracle.py against vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""

import os

import numpy as np
import pytest
import torch

from oracle import unitspeech_oracle as O

CASES = ["tiny_cfg", "tiny_nocfg", "tiny_textonly", "tiny_spkonly", "full_cfg",
         "d64_cfg", "d64_nocfg", "d64_textonly", "d64_spkonly", "full_nocfg10"]


def _load(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    dim, B, T, n = (int(v) for v in g["meta"])
    mults = tuple(int(v) for v in g["mults"])
    lengths = tuple(int(v) for v in g["lengths"])
    tg, sg, s = (float(v) for v in g["scales"])
    return g, dim, mults, B, T, n, lengths, tg, sg, s


@pytest.mark.parametrize("name", CASES)
def test_reverse_diffusion_matches_reference(golden_dir, name):
    g, dim, mults, B, T, n, lengths, tg, sg, s = _load(golden_dir, name)
    p = O.harness_params(dim=dim, dim_mults=mults, seed=1234, out_scale=s)
    z, mask, cond, spk, noise = O.harness_inputs(B, T, n, seed=11, scale=s, lengths=lengths)
    out = O.reverse_diffusion(p, z, mask, cond, spk, n, tg, sg, noise=noise, dim=dim, dim_mults=mults)
    ref = torch.from_numpy(g["out"])
    scale = float(ref.abs().max())
    # fp32 re-association only (batched vs per-utterance convs, closed-form-free): relative 2e-5
    assert float((out - ref).abs().max()) <= 2e-5 * max(scale, 1.0)


@pytest.mark.parametrize("name", CASES)
def test_estimator_matches_reference(golden_dir, name):
    g, dim, mults, B, T, n, lengths, tg, sg, s = _load(golden_dir, name)
    p = O.harness_params(dim=dim, dim_mults=mults, seed=1234, out_scale=s)
    z, mask, cond, spk, noise = O.harness_inputs(B, T, n, seed=11, scale=s, lengths=lengths)
    est = O.estimator_forward(p, z, mask, cond, torch.full((B,), 0.37), spk, dim, mults)
    ref = torch.from_numpy(g["est"])
    assert float((est - ref).abs().max()) <= 2e-5 * max(float(ref.abs().max()), 1e-3)


@pytest.mark.parametrize("n", [2, 4, 50])
def test_schedule_tables_bit_exact(golden_dir, n):
    g = np.load(os.path.join(golden_dir, "schedule.npz"))
    tb = O.schedule_tables(n, 0.05, 20.0)
    for k, v in tb.items():
        assert np.array_equal(v.numpy(), g[f"n{n}_{k}"]), k


def test_posemb_known_answer(golden_dir):
    g = np.load(os.path.join(golden_dir, "schedule.npz"))
    e = O.sinusoidal_pos_emb(torch.from_numpy(g["posemb_t"]), 128, 1000)
    assert np.array_equal(e.numpy(), g["posemb_128"])
    # SURVEY §8 c3 known answer for t = 0.99
    np.testing.assert_allclose(e[0, :3].numpy(), [-0.38786501, 0.74154478, -0.67208272], atol=2e-6)


def test_step_coefficients_known_answers():
    """SURVEY Appendix A.4 table (n=50)."""
    c = O.step_coefficients(50, 0.05, 20.0)
    np.testing.assert_allclose(c[0].numpy(), [1.2165391, 0.3945352, 0.5694744], rtol=2e-6)
    np.testing.assert_allclose(c[24].numpy(), [1.1054473, 0.2008360, 0.4225062], rtol=2e-6)
    np.testing.assert_allclose(c[49].numpy(), [1.0007490, 0.0014975, 0.0], rtol=2e-5, atol=1e-9)


def test_closed_form_step_equals_op_for_op():
    tb = O.schedule_tables(50, 0.05, 20.0)
    c = O.step_coefficients(50, 0.05, 20.0)
    g = torch.Generator().manual_seed(3)
    x, s, nz = (torch.randn(2, 80, 16, generator=g) for _ in range(3))
    m = torch.ones(2, 1, 16)
    for i in (0, 10, 49):
        ref = O.sampler_step(tb, 49 - i, x, s, nz, m)
        got = (c[i, 0] * x + c[i, 1] * s + c[i, 2] * nz) * m
        assert float((ref - got).abs().max()) < 5e-6


def test_n_timesteps_one_rejected():
    with pytest.raises(ValueError):
        O.schedule_tables(1, 0.05, 20.0)


@pytest.mark.parametrize("name", ["loss_d64", "loss_full"])
def test_oracle_loss_t_matches_reference(golden_dir, name):
    """oracle.loss_t / forward_diffusion vs UnitSpeech.loss_t of the unmodified reference (make_golden_loss.py)."""
    import sys
    sys.path.insert(0, golden_dir)
    from make_golden_loss import LOSS_CASES, loss_inputs
    dim, mults, B, T, lengths, ts, s = LOSS_CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    params = O.harness_params(dim=dim, dim_mults=mults, seed=1234, out_scale=s)
    x0, mask, cond, spk = loss_inputs(B, T, lengths)
    torch.manual_seed(77)
    z = torch.randn(x0.shape)
    loss, xt = O.loss_t(params, x0, mask, cond, torch.tensor(ts), spk, z, dim=dim, dim_mults=mults)
    assert np.abs(xt.numpy() - g["xt"]).max() <= 1e-6
    assert abs(float(loss) - float(g["loss"])) <= 2e-5 * float(g["loss"])
