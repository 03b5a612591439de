"""Sums an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name.
usage: python scripts/parse_train_launches.py launches.csv [skip_first_n]"""
import csv
import collections
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r and (r[0].isdigit() or r[0] == "ID")]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot = collections.defaultdict(lambda: [0, 0.0])
n = 0
for r in rows[1:]:
    n += 1
    if n <= skip:
        continue
    v = float(r[vi].replace(",", ""))
    us = v / 1e3 if r[ui] in ("ns", "nsecond") else (v if r[ui] in ("us", "usecond") else v * 1e3)
    name = r[ki].split("(")[0].split("<")[0]
    tot[name][0] += 1
    tot[name][1] += us
total = sum(v[1] for v in tot.values())
print(f"{n - skip} launches, {total / 1e3:.3f} ms of kernel time")
for k, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{us / 1e3:9.3f} ms {100 * us / total:5.1f}%  {c:5d} x {us / c:8.1f} us  {k}")
