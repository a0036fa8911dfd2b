"""Static code footprint of one kernel per source-line bucket (nvdisasm --print-line-info; needs -lineinfo).
usage: python scripts/sass_footprint.py file.o <substring of the mangled kernel name> [bucket_lines]"""
import re
import subprocess
import sys
import tempfile
import os
import glob

obj, pat = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 20
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cub = glob.glob(tmp + "/*.cubin")[0]
txt = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
sec, cur, counts = None, ("?", 0), {}
for ln in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        sec = m.group(1)
        continue
    if sec is None or pat not in sec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", ln):
        key = (sec[-40:], cur[0], cur[1] // bucket * bucket)
        counts[key] = counts.get(key, 0) + 1
tot = {}
for (s, f, b), n in counts.items():
    tot[s] = tot.get(s, 0) + n
for s, n in tot.items():
    print(f"{n:6d} SASS = {n * 16 / 1024:.1f} KiB  {s}")
for (s, f, b), n in sorted(counts.items()):
    if n >= 12:
        print(f"{n:6d}  {f}:{b}")
