"""Stall samples per reason and per named source-line range of one file (needs -lineinfo).
usage: python scripts/ncu_stall_regions.py X.ncu-rep file.cu name:lo-hi [name:lo-hi ...]"""
import csv
import io
import subprocess
import sys

rep, fname = sys.argv[1], sys.argv[2]
regions = []
for spec in sys.argv[3:]:
    name, rng = spec.split(":")
    lo, hi = (int(x) for x in rng.split("-"))
    regions.append((name, lo, hi))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True, errors="replace").stdout
cur, hdr, line = None, None, None
acc = {}
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
        continue
    if r[0] == "Line No":
        hdr = r
        stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        ix_ex = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "Function Name":
        continue
    if r[0]:
        line = int(r[0])
    if line is None or not r[2]:
        continue
    key = "other:" + cur
    if cur == fname:
        for name, lo, hi in regions:
            if lo <= line <= hi:
                key = name
                break
    d = acc.setdefault(key, {})
    try:
        d["inst"] = d.get("inst", 0.0) + float(r[ix_ex] or 0)
        for i, h in stalls:
            d[h] = d.get(h, 0.0) + float(r[i] or 0)
    except ValueError:
        pass
tot = sum(v for d in acc.values() for k, v in d.items() if k != "inst") or 1.0
toti = sum(d.get("inst", 0) for d in acc.values()) or 1.0
cols = ["stall_no_inst", "stall_wait", "stall_long_sb", "stall_short_sb", "stall_not_selected", "stall_selected",
        "stall_branch_resolving", "stall_math"]
print(f"{'region':28s} inst%  smpl% | " + " ".join(c[6:12].rjust(6) for c in cols) + "   (per cent of all samples)")
for key, d in sorted(acc.items(), key=lambda kv: -sum(v for k, v in kv[1].items() if k != "inst")):
    s = sum(v for k, v in d.items() if k != "inst")
    print(f"{key:28s} {100 * d.get('inst', 0) / toti:5.1f} {100 * s / tot:6.1f} | " +
          " ".join(f"{100 * d.get(c, 0) / tot:6.2f}" for c in cols))
