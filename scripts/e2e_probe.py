"""Wall-clock probe of one sampler call with a synchronise on both sides: device-tensor entry vs host-tensor entry."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from unitspeech_b200 import UnitSpeech
from unitspeech_b200.synthetic import random_init_state_dict, synthetic_inputs

B, T, n = int(sys.argv[1]) if len(sys.argv) > 1 else 1, int(sys.argv[2]) if len(sys.argv) > 2 else 256, 50
dec = UnitSpeech(80, 128, (1, 2, 4, 8), spk_emb_dim=256)
dec.load_state_dict(random_init_state_dict(dec))
dec = dec.cuda().eval()
host = [t.pin_memory() for t in synthetic_inputs(B, T, n)]
devt = [t.cuda() for t in host]


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return [round(x, 1) for x in ts]


z, m, c, s, nz = devt
print("device entry ms:", timed(lambda: dec(z, m, c, s, n, 1.0, 1.0, noise=nz)))
hz, hm, hc, hs, hn = host
print("host entry ms:  ", timed(lambda: dec(hz, hm, hc, hs, n, 1.0, 1.0, noise=hn)))
t0 = time.perf_counter()
for _ in range(3):
    dec(z, m, c, s, n, 1.0, 1.0, noise=nz)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("3 device calls: enqueue %.1f ms, drain %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
