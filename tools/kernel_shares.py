"""Per-kernel share of ONE registration from the ncu launch list of the bench command:
    python tools/kernel_shares.py gpurun_out/<tag>/launches_bench.csv [launches per registration] [out.csv of the last registration]"""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
rows = []
with open(path) as fh:
    lines = [ln for ln in fh if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"]) / 1e3))
per = int(sys.argv[2]) if len(sys.argv) > 2 else 607
nreg = len(rows) // per
last = rows[(nreg - 1) * per: nreg * per]
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as fh:
        fh.write("kernel,us\n")
        for k, t in last:
            fh.write('"%s",%.3f\n' % (k, t))
agg = OrderedDict()
for k, t in last:
    name = re.sub(r"\(.*", "", k).strip()
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches listed = {nreg} registrations of {per}; the last one:")
print(f"{'kernel':40s} {'launches':>8s} {'total us':>10s} {'share':>7s}")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:40]:40s} {n:8d} {t:10.1f} {100 * t / tot:6.1f}%")
print(f"{'sum':40s} {len(last):8d} {tot:10.1f}")
