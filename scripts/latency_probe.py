"""Latency of one small call (the reference callers' pattern, inference.py:128) under the four combinations of the two
small-call mechanisms: CUDA-graph replay of the sampler step and split-K of the few-tile convolutions.

    python scripts/latency_probe.py [--batch 1] [--frames 256] [--passes 5]

Prints one JSON line per combination: ms per 50-step text+speaker CFG pass (CUDA events, inputs resident)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from unitspeech_b200 import UnitSpeech  # noqa: E402
from unitspeech_b200.synthetic import random_init_state_dict, synthetic_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--frames", type=int, default=256)
ap.add_argument("--passes", type=int, default=5)
a = ap.parse_args()
dec = UnitSpeech(80, 128, (1, 2, 4, 8), spk_emb_dim=256)
dec.load_state_dict(random_init_state_dict(dec, out_scale=1 / 512))
dec = dec.cuda().eval()
n = 50
z, mask, cond, spk, noise = (t.cuda() for t in synthetic_inputs(a.batch, a.frames, n, seed=100))
ref = None
for graph, splitk in ((0, 0), (1, 0), (0, 1), (1, 1)):
    dec.graph_mode, dec.splitk_mode = graph, splitk
    for _ in range(2):
        out = dec(z, mask, cond, spk, n, 1.0, 1.0, noise=noise)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.passes):
        out = dec(z, mask, cond, spk, n, 1.0, 1.0, noise=noise)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.passes
    if ref is None:
        ref = out.clone()
    print(json.dumps({"batch": a.batch, "frames": a.frames, "graph": graph, "splitk": splitk, "ms_per_pass": round(ms, 3),
                      "frames_per_s": round(a.batch * a.frames / ms * 1e3, 1), "rtf": round(ms / 1e3 / (a.batch * a.frames * 256 / 22050), 5),
                      "max_abs_vs_eager_nosplit": float((out - ref).abs().max()), "finite": bool(torch.isfinite(out).all())}))
