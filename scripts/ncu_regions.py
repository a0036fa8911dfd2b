"""Share of executed warp instructions per named source-line range of one file (needs -lineinfo).
usage: python scripts/ncu_regions.py X.ncu-rep file.cu name:lo-hi [name:lo-hi ...]"""
import csv
import io
import subprocess
import sys

rep, fname = sys.argv[1], sys.argv[2]
ranges = [tuple(x.split(":")) for x in sys.argv[3:]]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True, errors="replace").stdout
cur, hdr, last, acc = None, None, None, {}
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr, ie = r, r.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0]:
        last = (cur, int(r[0]))
    if last is None or not r[2]:
        continue
    try:
        ex = float(r[ie] or 0)
    except ValueError:
        continue
    acc[last] = acc.get(last, 0) + ex
tot = sum(acc.values())
print(f"(shares of {tot:.4g} line-attributed warp instructions; inlined code is attributed to its own file)")
for name, rg in ranges:
    f = fname
    if "@" in name:
        name, f = name.split("@")
    lo, hi = map(int, rg.split("-"))
    s = sum(v for (ff, l), v in acc.items() if ff == f and lo <= l <= hi)
    print(f"{name:14s} {f}:{lo}-{hi:<5d} {100 * s / tot:6.2f}%")
others = {}
for (f, l), v in acc.items():
    others[f] = others.get(f, 0) + v
for f, v in sorted(others.items(), key=lambda kv: -kv[1])[:8]:
    print(f"file {f:30s} {100 * v / tot:6.2f}%")
