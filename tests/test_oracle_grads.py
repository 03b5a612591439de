"""The oracle's restatement of the fine-tune step (loss_t -> backward -> clip_grad_norm_ -> Adam) against gradients and
Adam trajectories of the unmodified reference (tests/golden/grads_*.npz, made by tests/golden/make_golden_grads.py)."""

import os

import numpy as np
import pytest
import torch

from oracle import unitspeech_oracle as O
from train_cases import CASES, FULL_KEYS, case_inputs, projection, reference_z

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["grads_d64", "grads_full"])
def test_oracle_gradients_and_adam_match_reference(name):
    dim, mults, B, T, lengths, ts, s, lr, K = CASES[name]
    if name == "grads_full":
        K = 2          # keeps the CPU suite short; the first two iterations pin the optimizer restatement
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    p0 = O.harness_params(dim=dim, dim_mults=mults, seed=1234, out_scale=s)
    p = dict(p0)
    x0, mask, cond, spk = case_inputs(B, T, lengths)
    t = torch.tensor(ts)
    state = {}
    for i in range(K):
        z = reference_z(x0.shape, i)
        loss, grads = O.loss_t_grads(p, x0, mask, cond, t, spk, z, dim=dim, dim_mults=mults)
        assert loss == pytest.approx(float(gold["losses"][i]), rel=2e-4)
        if i == 0:
            for k, g in grads.items():
                gn, gp = float(gold["gn/" + k]), float(gold["gp/" + k])
                assert float(g.double().norm()) == pytest.approx(gn, rel=1e-3, abs=1e-7), k
                assert float((g.double() * projection(k, g.shape).double()).sum()) == pytest.approx(gp, rel=5e-3, abs=1e-4 * gn + 1e-7), k
            for k in FULL_KEYS:
                if "gf/" + k in gold.files:
                    ref = torch.from_numpy(gold["gf/" + k])
                    assert (grads[k] - ref).norm() <= 1e-3 * ref.norm() + 1e-8, k
        p, total = O.clip_and_adam(p, grads, state, i + 1, lr=lr)
        assert total == pytest.approx(float(gold["norms"][i]), rel=1e-3)
    if K == CASES[name][8]:
        for k in p:
            assert float((p[k].double() - p0[k].double()).norm()) == pytest.approx(float(gold["dw/" + k]), rel=2e-2, abs=1e-7), k
