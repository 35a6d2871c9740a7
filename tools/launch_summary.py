#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
    python tools/launch_summary.py launches.csv[.gz] > profiles/rNN_launches_summary.txt"""
import csv, gzip, re, sys
from collections import defaultdict

path = sys.argv[1]
fh = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = [r for r in csv.reader(l for l in fh if l.startswith('"')) if len(r) > 14]
hdr, rows = rows[0], rows[1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    name = re.sub(r"\(.*", "", r[ki])            # drop the argument list
    name = re.sub(r"<(\d+), (false|true)>", r"<\1>", name)
    ns = float(r[vi].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[ui], 1.0)
    tot[name] += ns
    cnt[name] += 1
total = sum(tot.values())
print("ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --profile-only  (one S2ST step, batch 32x10s)")
print("per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
print("total kernel time %.1f ms over %d launches\n" % (total / 1e6, len(rows)))
print("%-60s %8s %10s %7s %9s" % ("kernel", "launches", "total_ms", "share", "avg_us"))
for name in sorted(tot, key=lambda n: -tot[n]):
    print("%-60s %8d %10.2f %6.1f%% %9.1f" % (name[:60], cnt[name], tot[name] / 1e6, 100 * tot[name] / total, tot[name] / cnt[name] / 1e3))
