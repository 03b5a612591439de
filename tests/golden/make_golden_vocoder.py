"""Generates tests/golden/vocoder_*.npz by running the UNMODIFIED reference BigVGAN generator
(unitspeech/vocoder/models.py) on seeded weights/inputs.  Build container only (needs /root/reference)."""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bigvgan_oracle as V  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CONFIGS = {
    # reduced widths (fast), every structural feature of the public config: 2 upsample rates incl. 4, 3 kernel sizes
    "vocoder_small": (dict(V.PUBLIC_22KHZ_80BAND, upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4],
                           upsample_initial_channel=128), 2, 12),
    "vocoder_small_snake_amp2": (dict(V.PUBLIC_22KHZ_80BAND, upsample_rates=[2, 2], upsample_kernel_sizes=[4, 4],
                                      upsample_initial_channel=64, resblock="2", activation="snake",
                                      snake_logscale=False, resblock_dilation_sizes=[[1, 3], [1, 3], [1, 3]]), 1, 9),
    "vocoder_public": (dict(V.PUBLIC_22KHZ_80BAND), 1, 6),
}


def main():
    ref_shim.load_reference()
    from unitspeech.vocoder.env import AttrDict
    from unitspeech.vocoder.models import BigVGAN
    torch.set_num_threads(8)
    for name, (h, B, T) in CONFIGS.items():
        gen = BigVGAN(AttrDict(h))
        gen.remove_weight_norm()
        p = V.harness_params(h)
        sd = gen.state_dict()
        # buffers (the kaiser-sinc filters) stay as constructed by the reference; everything else is seeded
        missing = [k for k in sd if k not in p and not k.endswith("filter")]
        assert not missing, missing
        sd.update(p)
        gen.load_state_dict(sd, strict=True)
        gen.eval()
        g = torch.Generator().manual_seed(17)
        mel = torch.randn(B, h["num_mels"], T, generator=g) * 2 - 4      # log-mel-like range
        with torch.no_grad():
            out = gen(mel)
        filt = sd["activation_post.upsample.filter"].reshape(-1)
        np.savez(os.path.join(OUT, name + ".npz"), out=out.numpy(), filt=filt.numpy(), B=B, T=T)
        print(name, out.shape, float(out.abs().max()), sum(v.numel() for k, v in p.items()))


if __name__ == "__main__":
    main()
