"""Aggregate an `ncu --page source --csv` dump: stall-reason totals and the hottest SASS regions.
usage: ncu -i X.ncu-rep --page source --csv > /tmp/src.csv ; python scripts/ncu_source_summary.py /tmp/src.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
body = [r for r in rows[2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: sum(float(r[ix[s]] or 0) for r in body) for s in stalls}
allsamp = sum(tot.values())
print("stall reasons (all samples):")
for s, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"  {s:28s} {v:10.0f} {100 * v / allsamp:6.2f}%")
inst = sum(float(r[ix["Instructions Executed"]] or 0) for r in body)
print(f"total warp instructions {inst:.4g}, SASS lines {len(body)}")
# opcode histogram weighted by executed instructions
ops = {}
for r in body:
    op = r[ix["Source"]].split()[0] if r[ix["Source"]] else "?"
    if op.startswith("@"):
        op = r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    ops[op] = ops.get(op, 0) + float(r[ix["Instructions Executed"]] or 0)
print("opcode mix (executed warp instructions):")
for op, v in sorted(ops.items(), key=lambda kv: -kv[1])[:25]:
    print(f"  {op:12s} {100 * v / inst:6.2f}%")
# hottest lines by samples
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print(f"top {n} SASS lines by stall samples:")
for r in sorted(body, key=lambda r: -float(r[ix['# Samples']] or 0))[:n]:
    top = max(stalls, key=lambda s: float(r[ix[s]] or 0))
    print(f"  {r[ix['Address']][-6:]} {float(r[ix['# Samples']]):8.0f} ex={float(r[ix['Instructions Executed']]):.3g} "
          f"thr={r[ix['Avg. Threads Executed']]:>5s} {top:18s} {r[ix['Source']][:70]}")
