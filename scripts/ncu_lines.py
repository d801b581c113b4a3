#!/usr/bin/env python
"""Join an ncu SASS-level source page with nvdisasm line info: per source line, the share of
warp-stall samples and executed instructions. Usage:
  scripts/ncu_lines.py <report.ncu-rep> <cubin> <kernel-substring> [top]
"""
import csv, re, subprocess, sys, collections, io

rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# map address -> (file,line) per function
addr2line = {}
cur_fn, cur_line, inl = None, None, None
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m:
        cur_fn = m.group(1); cur_line = None; continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        f = m.group(1).split("/")[-1]
        cur_line = (f, int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur_fn and kname in cur_fn:
        addr2line[int(m.group(1), 16)] = cur_line
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = csv.DictReader(io.StringIO("\n".join(lines[start:])))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
base = None
for row in rd:
    try:
        a = int(row["Address"], 16)
    except Exception:
        continue
    if base is None:
        base = a
    key = addr2line.get(a - base, ("?", 0))
    s = int(row["# Samples"] or 0); ie = int(row["Instructions Executed"] or 0); te = int(row["Thread Instructions Executed"] or 0)
    for i, v in enumerate((s, ie, te)):
        agg[key][i] += v; tot[i] += v
print(f"total samples {tot[0]}  warp-inst {tot[1]}  thread-inst {tot[2]}  avg active {tot[2]/max(tot[1],1):.1f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(k[0]):18s}:{k[1]:4d}  samples {100*v[0]/tot[0]:5.1f}%  winst {100*v[1]/tot[1]:5.1f}%  active {v[2]/max(v[1],1):5.1f}")
