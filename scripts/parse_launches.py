"""Summarises an ncu launch list (gpu__time_duration.sum CSV) of scripts/profile_pass.py: per-kernel totals and the
per-conv TFLOP/s of one diffusion step (conv order reconstructed from the plan for Be=48, T=512)."""
import collections, csv, sys
path = sys.argv[1]
args = [a for a in sys.argv[2:] if not a.startswith("-")]
Be, T = (int(args[0]) if len(args) > 0 else 48), (int(args[1]) if len(args) > 1 else 512)
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rows = [(r['Kernel Name'].split('(')[0].replace('usb::', ''), float(r['Metric Value']) / 1e3) for r in csv.DictReader(lines)]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in rows:
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot/1e3:.3f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]:
    print(f"  {k[:60]:60s} n={v[0]:4d} {v[1]/1e3:9.3f} ms {v[1]/tot:6.1%}")
C = [128, 256, 512, 1024]; H = [80, 40, 20, 10]; W = [T >> l for l in range(4)]
ops = []
def conv(name, l, cin, cout, taps, scale=1.0): ops.append((name, l, cin, cout, taps, 2 * Be * H[l] * W[l] * scale * cout * cin * taps))
def resnet(name, l, cin, cout, first=False):
    if not first: conv(name + '.c1', l, cin, cout, 9)
    conv(name + '.c2', l, cout, cout, 9)
    if cin != cout and not first: conv(name + '.res', l, cin, cout, 1)
def attn(name, l, c): conv(name + '.qkv', l, c, 384, 1); conv(name + '.out', l, 128, c, 1)
for k in range(4):
    resnet(f'd{k}.0', k, 2 if k == 0 else C[k - 1], C[k], first=(k == 0)); resnet(f'd{k}.1', k, C[k], C[k]); attn(f'd{k}.2', k, C[k])
    if k < 3: conv(f'd{k}.down', k, C[k], C[k], 9, 0.25)
resnet('mid1', 3, 1024, 1024); attn('mida', 3, 1024); resnet('mid2', 3, 1024, 1024)
for k in range(3):
    j = 3 - k
    resnet(f'u{k}.0', j, 2 * C[j], C[j - 1]); resnet(f'u{k}.1', j, C[j - 1], C[j - 1]); attn(f'u{k}.2', j, C[j - 1]); conv(f'u{k}.up', j, C[j - 1], C[j - 1], 16)
conv('final', 0, 128, 128, 9)
convs = [r for r in rows if r[0].startswith('conv_igemm')]
step = convs[-60:]
cls = collections.defaultdict(lambda: [0.0, 0.0])
for o, c in zip(ops, step):
    kind = '1x1' if o[4] == 1 else ('3x3/T N=128' if o[3] == 128 else '3x3/T N>=256')
    cls[kind][0] += c[1]; cls[kind][1] += o[5]
    if '-v' in sys.argv: print(f"  {o[0]:10s} l{o[1]} {o[2]:5d}->{o[3]:5d} taps {o[4]:2d} {c[1]:8.1f} us {o[5]/(c[1]*1e-6)/1e12:7.1f} TF")
for k, (us, fl) in cls.items(): print(f"  conv class {k:14s} {us/1e3:7.3f} ms/step {fl/(us*1e-6)/1e12:7.1f} TFLOP/s")
