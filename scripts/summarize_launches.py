"""Summarises an ncu launch list of scripts/profile_pass.py (one metric row per launch and metric):

    python scripts/summarize_launches.py <launches.csv> <batch> <frames> [-v] [--traffic-json out.json]

* per-kernel totals and shares of ONE diffusion step (the last complete step in the file),
* per-conv TFLOP/s, tensor-pipe active % and DRAM bytes against the algorithmic bytes (conv order reconstructed from the
  plan in csrc/engine.cu:build_plan for Be = 3 * batch rows, T = frames),
* with --traffic-json: the per-launch DRAM traffic of the conv class and of gn_apply, as quoted by bench.py's `roofline.traffic`.

Times under ncu are cold-cache and serialised: compare SHARES with bench.py's own CUDA-event profile, not absolutes."""
import collections
import csv
import json
import sys

args = [a for a in sys.argv[1:] if not a.startswith("-")]
path, batch, T = args[0], int(args[1]), int(args[2])
verbose = "-v" in sys.argv
traffic_out = sys.argv[sys.argv.index("--traffic-json") + 1] if "--traffic-json" in sys.argv else None
if traffic_out in args:
    args.remove(traffic_out)
Be = 3 * batch

with open(path) as f:
    lines = [ln for ln in f if ln.startswith('"')]
launches = collections.OrderedDict()
for r in csv.DictReader(lines):
    d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"].split("(")[0].replace("usb::", "").replace("void ", "")})
    try:
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        pass
    d.setdefault("unit:" + r["Metric Name"], r["Metric Unit"])
rows = list(launches.values())


def t_us(d):
    v, u = d.get("gpu__time_duration.sum", 0.0), d.get("unit:gpu__time_duration.sum", "ns")
    return v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)


def dram(d):
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        if m not in d:
            return None
        u = d.get("unit:" + m, "byte")
        tot += d[m] * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    return tot


# ---- the last complete step: from the last emb_combine_kernel to the final_kernel after it
idx_comb = [i for i, d in enumerate(rows) if d["name"].startswith("emb_combine")]
start = idx_comb[-1]
end = next(i for i in range(start, len(rows)) if rows[i]["name"].startswith("final_kernel"))
if end is None:
    start = idx_comb[-2]
    end = next(i for i in range(start, len(rows)) if rows[i]["name"].startswith("final_kernel"))
step = rows[start:end + 1]
agg = collections.defaultdict(lambda: [0, 0.0])
for d in step:
    agg[d["name"]][0] += 1
    agg[d["name"]][1] += t_us(d)
tot = sum(v[1] for v in agg.values())
print(f"one diffusion step at {batch} utt x {T} frames (Be = {Be}): {len(step)} launches, {tot / 1e3:.3f} ms of kernel time under ncu")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"  {k[:58]:58s} n={v[0]:3d} {v[1] / 1e3:9.3f} ms {v[1] / tot:6.1%}")

# ---- conv plan (same order as build_plan)
C = [128, 256, 512, 1024]
H = [80, 40, 20, 10]
W = [T >> l for l in range(4)]
ops = []


def conv(name, l, cin, cout, taps, scale=1.0, out_scale=1.0):
    px = Be * H[l] * W[l]
    flop = 2.0 * px * scale * cout * cin * taps
    # algorithmic bytes: every input and output activation once (fp16) + the weights once
    by = 2.0 * px * cin + 2.0 * px * scale * out_scale * cout + 2.0 * cout * cin * taps
    ops.append((name, l, cin, cout, taps, flop, by))


def resnet(name, l, cin, cout, first=False):
    if not first:
        conv(name + ".c1", l, cin, cout, 9)
    conv(name + ".c2", l, cout, cout, 9)
    if cin != cout and not first:
        conv(name + ".res", l, cin, cout, 1)


def attn(name, l, c):
    if c <= 256:      # fused-q form: k,v GEMM (256 outputs) + one per-sample CxC conv on x
        conv(name + ".kv", l, c, 256, 1)
        conv(name + ".out", l, c, c, 1)
    else:
        conv(name + ".qkv", l, c, 384, 1)
        conv(name + ".out", l, 128, c, 1)


for k in range(4):
    resnet(f"d{k}.0", k, 2 if k == 0 else C[k - 1], C[k], first=(k == 0))
    resnet(f"d{k}.1", k, C[k], C[k])
    attn(f"d{k}.2", k, C[k])
    if k < 3:
        conv(f"d{k}.down", k, C[k], C[k], 9, 0.25)
resnet("mid1", 3, 1024, 1024)
attn("mida", 3, 1024)
resnet("mid2", 3, 1024, 1024)
for k in range(3):
    j = 3 - k
    resnet(f"u{k}.0", j, 2 * C[j], C[j - 1])
    resnet(f"u{k}.1", j, C[j - 1], C[j - 1])
    attn(f"u{k}.2", j, C[j - 1])
    conv(f"u{k}.up", j, C[j - 1], C[j - 1], 16, 1.0, 4.0)       # 4 phases x 4 taps per input pixel, output 2H x 2W
conv("final", 0, 128, 128, 9)

convs = [d for d in step if d["name"].startswith("conv_igemm")]
assert len(convs) == len(ops), (len(convs), len(ops))
cls = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0.0, 0])
for o, d in zip(ops, convs):
    us = t_us(d)
    kind = "1x1" if o[4] == 1 else ("3x3 s1 halo (l0/l1)" if d["name"].startswith("conv_igemm_halo") else
                                    ("strided/transposed" if o[0].endswith(("down", "up")) else "3x3 s1 swapped (l2/l3)"))
    dr = dram(d)
    c = cls[kind]
    c[0] += us; c[1] += o[5]; c[2] += o[6]; c[3] += dr or 0.0; c[4] += 1
    if verbose:
        tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        clk = d.get("sm__cycles_elapsed.avg.per_second")
        print(f"  {o[0]:10s} l{o[1]} {o[2]:5d}->{o[3]:5d} taps {o[4]:2d} {d['name'][:26]:26s} {us:8.1f} us {o[5] / (us * 1e-6) / 1e12:7.1f} TF"
              + (f"  tensor {tp:5.1f}%" if tp is not None else "") + (f"  dram {dr / 1e6:8.1f} MB (alg {o[6] / 1e6:8.1f})" if dr is not None else "")
              + (f"  {clk / 1e9 if clk > 1e6 else clk:5.3f} GHz" if clk else ""))
print("conv classes of the step:")
for k, (us, fl, by, dr, n) in cls.items():
    print(f"  {k:24s} n={n:2d} {us / 1e3:7.3f} ms {fl / (us * 1e-6) / 1e12:7.1f} TFLOP/s" + (f"  dram/algorithmic bytes {dr / by:5.2f}" if dr else ""))
us_all = sum(c[0] for c in cls.values())
fl_all = sum(c[1] for c in cls.values())
print(f"  all {len(convs)} conv launches: {us_all / 1e3:.3f} ms, {fl_all / (us_all * 1e-6) / 1e12:.1f} TFLOP/s, share of step {us_all / tot:.1%}")

if traffic_out:
    gn = [d for d in step if d["name"].startswith("gn_apply")]
    rec = {
        "conv_igemm": {"dram_bytes_per_launch": sum(dram(d) for d in convs) / len(convs),
                       "algorithmic_bytes_per_launch": sum(o[6] for o in ops) / len(ops),
                       "launches": len(convs),
                       "note": "dram__bytes_read.sum + dram__bytes_write.sum averaged over the 60 conv launches of one evaluation "
                               "(ncu, cold cache, this workload); algorithmic = every input / output activation once + weights once, fp16"},
        "gn_apply": {"dram_bytes_per_launch": sum(dram(d) for d in gn) / len(gn), "launches": len(gn)},
        "source": path,
    }
    try:
        with open(traffic_out) as f:
            allrec = json.load(f)
    except Exception:  # noqa: BLE001
        allrec = {}
    allrec[f"{batch}x{T}"] = rec
    with open(traffic_out, "w") as f:
        json.dump(allrec, f, indent=1)
    print("wrote", traffic_out)
