"""Executed warp instructions and stall samples per CUDA source line from an .ncu-rep (needs -lineinfo).
usage: python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
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
tot = sum(v[0] for v in lines.values()) or 1.0
tots = sum(v[1] for v in lines.values()) or 1.0
print(f"total warp instructions {tot:.4g}, samples {tots:.4g}")
print(" inst%  smpl%  lanes  file:line  source")
for (f, ln, src), (ex, smp, thr) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * ex / tot:6.2f} {100 * smp / tots:6.2f} {thr / max(ex, 1):6.1f}  {f}:{ln}  {src}")
