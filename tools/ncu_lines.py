"""Join an `ncu --page source --csv` SASS dump with nvdisasm line info: executed instructions per CUDA source line.
usage: ncu_lines.py <src.csv> <kernel-mangled-substring> [top] [file-filter]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src_csv, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ffilter = sys.argv[4] if len(sys.argv) > 4 else None
lo, hi = (int(sys.argv[5]), int(sys.argv[6])) if len(sys.argv) > 6 else (0, 10 ** 9)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "drl-dronenavigation_b200", "libdronenav.so")], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith("\t.section\t.text.") and kname in l)
cur, a2l = None, {}
for l in dis[start + 1:]:
    if l.startswith("\t.section") and a2l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        a2l[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
h = rows[1]
ie, sm = h.index("Instructions Executed"), h.index("# Samples")
base = int(rows[2][0], 16)
agg, smp, tot = collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    n = int(r[ie]); tot += n
    k = a2l.get(int(r[0], 16) - base)
    agg[k] += n; smp[k] += int(r[sm])
nwarps = int(rows[2][ie])
print(f"total warp-instructions {tot}  per warp {tot / nwarps:.1f}  (warps {nwarps})")
srcs = {f: open(os.path.join(ROOT, "drl-dronenavigation_b200", "csrc", f)).read().split("\n") for f in ("dn_device.cuh", "dronenav.cu")}
sel = 0
for k, n in agg.most_common():
    if ffilter and (not k or k[0] != ffilter or not (lo <= k[1] <= hi)):
        continue
    sel += n
print(f"selected {sel / nwarps:.1f} inst/warp")
for k, n in agg.most_common(top):
    if ffilter and (not k or k[0] != ffilter or not (lo <= k[1] <= hi)):
        continue
    txt = srcs[k[0]][k[1] - 1].strip()[:100] if k and k[0] in srcs else ""
    print(f"{n / nwarps:7.1f} inst/warp smp={smp[k]:5d} {str(k):28s} {txt}")
