"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel markdown table."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
acc = defaultdict(lambda: [0, 0.0])
for r in rd:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*$", "", r[ki])
    name = name.replace("void ", "").replace("ncme::", "").replace("<unnamed>::", "")
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
    acc[name][0] += 1
    acc[name][1] += us
tot = sum(v[1] for v in acc.values())
print("| kernel | launches | total us | avg us | share of all captured |")
print("|---|---|---|---|---|")
for k, (c, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    if t / tot < 0.002:
        continue
    print(f"| `{k}` | {c} | {t:.1f} | {t/c:.2f} | {100*t/tot:.1f} % |")
