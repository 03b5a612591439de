"""Generates tests/golden/*.npz by running the UNMODIFIED upstream reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

Every vector is produced by the reference's own classes
(unitspeech/unitspeech.py: UnitSpeech, GradLogPEstimator2d, SinusoidalPosEmb)
with the seeded harness weights/inputs of oracle/unitspeech_oracle.py loaded
through ``load_state_dict(strict=True)`` — so the fixtures also pin the
state_dict key names and shapes.  Only outputs (and seeds) are stored; inputs
are regenerated from the seeds at test time.
"""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle import unitspeech_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (dim, dim_mults, B, T, lengths, n_steps, tg, sg, out_scale)
    "tiny_cfg": (16, (1, 2, 4, 8), 2, 16, (16, 11), 4, 1.0, 1.0, 1.0 / 8),
    "tiny_nocfg": (16, (1, 2, 4, 8), 1, 24, (24,), 3, 0.0, 0.0, 1.0 / 8),
    "tiny_textonly": (16, (1, 2), 1, 16, (13,), 3, 1.5, 0.0, 1.0 / 8),
    "tiny_spkonly": (16, (1, 2), 1, 16, (16,), 3, 0.0, 0.7, 1.0 / 8),
    "full_cfg": (128, (1, 2, 4, 8), 2, 32, (32, 27), 3, 1.0, 1.0, 1.0 / 512),
    # dim=64 cases: the smallest width the CUDA library supports (channels are walked in 64-wide K steps)
    "d64_cfg": (64, (1, 2), 2, 16, (16, 11), 4, 1.0, 1.0, 1.0 / 32),
    "d64_nocfg": (64, (1, 2, 4), 1, 24, (24,), 3, 0.0, 0.0, 1.0 / 32),
    "d64_textonly": (64, (1, 2), 1, 16, (13,), 3, 1.5, 0.0, 1.0 / 32),
    "d64_spkonly": (64, (1, 2), 2, 16, (16, 9), 3, 0.0, 0.7, 1.0 / 32),
    "full_nocfg10": (128, (1, 2, 4, 8), 1, 40, (37,), 10, 0.0, 0.0, 1.0 / 512),
}


def build_reference(U, dim, dim_mults, params):
    dec = U.UnitSpeech(n_feats=80, dim=dim, dim_mults=dim_mults, beta_min=0.05, beta_max=20, pe_scale=1000,
                       spk_emb_dim=256)
    missing = dec.load_state_dict(params, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return dec.eval()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    U = ref_shim.load_reference()

    # ---- schedule + positional-embedding known answers (n = 50, the headline setting) ----
    dec = U.UnitSpeech(n_feats=80, dim=16, dim_mults=(1, 2), spk_emb_dim=256)
    sched = {}
    for n in (2, 4, 50):
        z = torch.zeros(1, 80, 8)
        # run the reference's own table construction (it happens inside reverse_diffusion);
        # n steps of a dim-16 net on T=8 is cheap
        p = O.harness_params(dim=16, dim_mults=(1, 2), seed=7)
        dec.load_state_dict(p, strict=True)
        dec.reverse_diffusion(z, torch.ones(1, 1, 8), z, torch.zeros(1, 1, 256), n)
        for name in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_one_minus_alphas_cumprod",
                     "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance"):
            sched[f"n{n}_{name}"] = getattr(dec, name).numpy().copy()
    t = torch.tensor([0.99, 0.51, 0.01])
    sched["posemb_t"] = t.numpy()
    sched["posemb_128"] = U.SinusoidalPosEmb(128)(t, scale=1000).numpy()
    np.savez(os.path.join(OUT, "schedule.npz"), **sched)

    # ---- sampler end-to-end cases ----
    for name, (dim, mults, B, T, lengths, n, tg, sg, s) in CASES.items():
        params = O.harness_params(dim=dim, dim_mults=mults, seed=1234, out_scale=s)
        z, mask, cond, spk, noise = O.harness_inputs(B, T, n, seed=11, scale=s, lengths=lengths)
        dec = build_reference(U, dim, mults, params)
        out = ref_shim.run_reference_per_utterance(dec, z, mask, cond, spk, noise, n, tg, sg)
        # one estimator evaluation at t=0.37 with the same inputs (B>1 is fine for the estimator itself)
        tt = torch.full((B,), 0.37)
        with torch.no_grad():
            est = dec.estimator(z, mask, cond, tt, spk)
        np.savez(os.path.join(OUT, f"{name}.npz"), out=out.numpy(), est=est.numpy(),
                 meta=np.array([dim, B, T, n], dtype=np.int64), mults=np.array(mults, dtype=np.int64),
                 lengths=np.array(lengths, dtype=np.int64), scales=np.array([tg, sg, s], dtype=np.float64))
        print(name, "out absmax", float(out.abs().max()), "est absmax", float(est.abs().max()))


if __name__ == "__main__":
    main()
