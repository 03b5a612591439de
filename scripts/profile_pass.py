"""Short profiling workload: a bench configuration (--batch x --frames, text+speaker CFG) for a few diffusion
steps, once as warm-up and once measured.  Used under ncu (launch list / --set full); prints nothing timed."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from unitspeech_b200 import UnitSpeech  # noqa: E402
from unitspeech_b200.synthetic import random_init_state_dict, synthetic_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--frames", type=int, default=512)
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--graph", type=int, default=0, help="sampler CUDA-graph mode: -1 auto, 0 off (plain launches for ncu), 1 on")
a = ap.parse_args()
dec = UnitSpeech(80, 128, (1, 2, 4, 8), spk_emb_dim=256)
dec.load_state_dict(random_init_state_dict(dec, out_scale=1 / 512))
dec = dec.cuda().eval()
dec.graph_mode = a.graph
z, mask, cond, spk, noise = (t.cuda() for t in synthetic_inputs(a.batch, a.frames, a.steps, seed=100))
for _ in range(a.passes):
    l0 = dec.launch_count
    out = dec(z, mask, cond, spk, a.steps, 1.0, 1.0, noise=noise)
    torch.cuda.synchronize()
    print("launches in pass:", dec.launch_count - l0, "finite:", bool(torch.isfinite(out).all()))
