"""The product-side synthetic generators (bench/smoke) produce the same tensors as the parity harness in oracle/."""

import torch

from oracle import unitspeech_oracle as O
from unitspeech_b200 import UnitSpeech
from unitspeech_b200.synthetic import random_init_state_dict, synthetic_inputs


def test_random_init_matches_harness_params():
    for dim, mults in ((64, (1, 2)), (128, (1, 2, 4, 8))):
        dec = UnitSpeech(80, dim, mults, spk_emb_dim=256)
        a = random_init_state_dict(dec, seed=1234, out_scale=1 / 512)
        b = O.harness_params(dim=dim, dim_mults=mults, seed=1234, out_scale=1 / 512)
        assert list(a) == list(b)
        assert all(torch.equal(a[k], b[k]) for k in a)


def test_synthetic_inputs_match_harness_inputs():
    a = synthetic_inputs(2, 16, 3, seed=5, scale=1 / 32, lengths=(16, 9))
    b = O.harness_inputs(2, 16, 3, seed=5, scale=1 / 32, lengths=(16, 9))
    assert all(torch.equal(x, y) for x, y in zip(a, b))
