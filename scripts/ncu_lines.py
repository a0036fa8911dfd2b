"""Top CUDA source lines by executed warp instructions / stall samples from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > f.csv`.  usage: ncu_lines.py f.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname, hdr, out = "?", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        i_ex, i_s, i_thr = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
    elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
        try:
            out.append((float(r[i_ex] or 0), float(r[i_s] or 0), fname, int(r[0]), r[1].strip()[:86], r[i_thr]))
        except ValueError:
            pass
tot = sum(o[0] for o in out)
tots = sum(o[1] for o in out)
print(f"total warp instructions {tot:.4g}, samples {tots:.0f}")
for ex, s, f, ln, src, thr in sorted(out, key=lambda o: -o[0])[:n]:
    print(f"{100 * ex / tot:5.1f}% inst {100 * s / max(tots, 1):5.1f}% smp thr={thr:>5s} {f}:{ln:<4d} {src}")
