#!/usr/bin/env python
"""Top stall-sample SASS lines of an ncu report's source page: python tools/ncu_hot.py file.ncu-rep [N]"""
import csv, subprocess, sys
rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
si = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in body)
print("total samples", tot)
for idx, r in sorted(enumerate(body), key=lambda x: -int(x[1][si]))[:n]:
    st = sorted(((int(r[i]), hdr[i]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:3]
    print(f"{int(r[si]):7d} {100.0*int(r[si])/tot:5.1f}%  line {idx:5d}  {r[1].strip()[:70]:70s} {st}")
