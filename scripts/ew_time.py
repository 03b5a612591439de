"""Tuning aid: CUDA-event timing of the streaming kernels at bench shapes via the profiled reverse-diffusion pass."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from unitspeech_b200 import UnitSpeech
from unitspeech_b200.synthetic import random_init_state_dict, synthetic_inputs
dec = UnitSpeech(80, 128, (1, 2, 4, 8), spk_emb_dim=256)
dec.load_state_dict(random_init_state_dict(dec, out_scale=1 / 512)); dec = dec.cuda().eval()
n = 4
z, mask, cond, spk, noise = (t.cuda() for t in synthetic_inputs(16, 512, n, seed=100))
for _ in range(2): dec(z, mask, cond, spk, n, 1.0, 1.0, noise=noise)
torch.cuda.synchronize()
dec.set_profiling(True)
dec(z, mask, cond, spk, n, 1.0, 1.0, noise=noise)
prof = dec.get_profile(); dec.set_profiling(False)
for k, (ms, work, cnt) in prof.items():
    print(f"{k:12s} {ms / n:8.3f} ms/step  {cnt // n:4d} launches/step  {work / ms / 1e6 if ms else 0:10.1f} {'GFLOP/s' if k == 'conv_igemm' else 'GB/s'}")
print("total ms/step", sum(v[0] for v in prof.values()) / n)
