#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel: joins the SASS page of an ncu report (instructions executed,
stall samples per SASS instruction) with the line table of the cubin (nvdisasm -g).
usage: tools/ncu_lines.py <report.ncu-rep> <cubin | object file> <mangled kernel name> [top N]
(the object must be the build the report was taken from)"""
import collections
import csv
import re
import subprocess
import sys

rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
for k in range(2, len(rows)):            # several captured launches: keep the first one
    if rows[k] and rows[k][0] == "Kernel Name":
        rows = rows[:k]
        break
hdr = rows[1]
ie, iss, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
prof = [(r[isrc].strip(), int(r[ie]), int(r[iss])) for r in rows[2:] if len(r) > ie]
if cubin.endswith(".o") or cubin.endswith(".so"):        # a host object: take the sm_100a cubin out of it first
    import glob, os, tempfile
    td = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(cubin)], cwd=td, capture_output=True, text=True)
    cands = sorted(glob.glob(os.path.join(td, "*.cubin")), key=os.path.getsize)
    assert cands, "no cubin in " + cubin
    cubin = cands[-1]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], None, False
for l in dis:
    if l.startswith(".text."):
        on = l.startswith(".text." + kern + ":")
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
assert len(lines) == len(prof), (len(lines), len(prof))
tot = sum(p[1] for p in prof)
tots = sum(p[2] for p in prof)
agg = collections.defaultdict(lambda: [0, 0])
for ln, (s, e, sm) in zip(lines, prof):
    agg[ln][0] += e
    agg[ln][1] += sm
print(f"total warp instructions {tot}, samples {tots}")
for ln, (e, sm) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{ln[0]:18s} {ln[1]:5d}  {e / tot * 100:6.2f}% inst  {sm / tots * 100:6.2f}% samples")
