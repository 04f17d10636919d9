#!/usr/bin/env python
"""Summarise an ncu report's source page: stall-reason mix and the hottest SASS instructions per kernel.
usage: python tools/ncu_src.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
seen = set()
for si, s in enumerate(secs):
    name = rows[s][1][:90]
    if name in seen:
        continue
    seen.add(name)
    hdr = rows[s + 1]
    body = rows[s + 2: secs[si + 1] if si + 1 < len(secs) else len(rows)]
    stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = {h: 0.0 for _, h in stall}
    for r in body:
        for i, h in stall:
            try:
                tot[h] += float(r[i])
            except (ValueError, IndexError):
                pass
    al = sum(tot.values()) or 1.0
    print("==", name)
    print("   stalls %:", {h[6:]: round(100 * v / al, 1) for h, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v / al > 0.01})
    ns, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    tots = sum(float(r[ns] or 0) for r in body) or 1.0
    print("   instructions executed (warp):", sum(float(r[ie] or 0) for r in body))
    for r in sorted(body, key=lambda r: -float(r[ns] or 0))[:top_n]:
        why = {h[6:]: r[i] for i, h in stall if r[i] not in ("0", "")}
        top = sorted(why.items(), key=lambda kv: -float(kv[1]))[:2]
        print("   %5.1f%%  %-70s %s" % (100 * float(r[ns] or 0) / tots, r[src][:70], top))
