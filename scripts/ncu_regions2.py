"""Executed warp instructions, stall samples and active lanes per named source-line REGION of an .ncu-rep (-lineinfo).
usage: python scripts/ncu_regions2.py REP file:lo-hi=name [...]   (uncovered lines are reported as other:<file>)"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
specs = sys.argv[2:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True, errors="replace").stdout
cur, hdr, lines = None, None, {}
last_key = None
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        hdr = None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ix_ex = hdr.index("Instructions Executed")
        ix_smp = hdr.index("# Samples")
        ix_thr = hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0]:
        last_key = (cur, int(r[0]), r[1].strip()[:100])
    if last_key is None or not r[2]:
        continue
    try:
        ex, smp, thr = float(r[ix_ex] or 0), float(r[ix_smp] or 0), float(r[ix_thr] or 0)
    except ValueError:
        continue
    d = lines.setdefault(last_key, [0.0, 0.0, 0.0])
    d[0] += ex
    d[1] += smp
    d[2] += thr

regions = []
for spec in specs:
    rng, name = spec.split("=")
    f, lh = rng.split(":")
    lo, hi = lh.split("-")
    regions.append((f, int(lo), int(hi), name))
acc = {}
for (f, ln, src), (ex, smp, thr) in lines.items():
    name = "other:" + str(f)
    for rf, lo, hi, nm in regions:
        if f == rf and lo <= ln <= hi:
            name = nm
            break
    d = acc.setdefault(name, [0.0, 0.0, 0.0])
    d[0] += ex; d[1] += smp; d[2] += thr
tot = sum(v[0] for v in acc.values()) or 1.0
tots = sum(v[1] for v in acc.values()) or 1.0
print(f"total warp instructions {tot:.4g}, samples {tots:.4g}")
for name, (ex, smp, thr) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print(f"  {name:32s} {100 * ex / tot:6.2f}% inst {100 * smp / tots:6.2f}% samples  lanes {thr / max(ex, 1):5.1f}")
