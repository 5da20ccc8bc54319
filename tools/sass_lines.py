#!/usr/bin/env python
"""Attribute the per-instruction counters of an ncu SASS source page to CUDA source lines.

    python tools/sass_lines.py <report.ncu-rep> <kernel name> <cubin built from the same source> [top]

Joins `ncu --page source --csv` (per-SASS-instruction "Instructions Executed" and stall samples) with the
line table of `nvdisasm -g` by instruction offset.  Development tool.
"""
import csv, io, re, subprocess, sys, collections

rep, kern, cubin = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ie, samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = rows[hdr_i + 1:]
# a report may hold several launches of the kernel back to back: keep the first
first_addr = int(data[0][0], 16)
insts = []
for r in data:
    if len(r) <= ie:
        break
    off = int(r[0], 16) - first_addr
    if insts and off == 0:
        break
    insts.append((off, r[1].strip(), int(r[ie] or 0), int(r[samp] or 0)))
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of, cur, inside = {}, None, False
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
agg = collections.defaultdict(lambda: [0, 0, 0])
tot_i = sum(i[2] for i in insts)
tot_s = sum(i[3] for i in insts)
for off, txt, n, s in insts:
    k = line_of.get(off, ("?", 0))
    agg[k][0] += n; agg[k][1] += s; agg[k][2] += 1
print(f"{kern}: {tot_i:,} warp instructions, {tot_s:,} samples, {len(insts)} SASS instructions")
for k, (n, s, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n / max(tot_i, 1) * 100:5.1f}% inst {s / max(tot_s, 1) * 100:5.1f}% samples  {k[0]}:{k[1]}  ({c} SASS)")
