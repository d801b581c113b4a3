#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (nvdisasm -g line info).
Usage: scripts/sass_lines.py <cubin> <kernel-substring> [top]"""
import re, subprocess, sys, collections
cubin, kname = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
cur_fn, cur_line = None, None
cnt = collections.Counter()
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m:
        cur_fn = m.group(1); cur_line = None; continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur_fn and kname in cur_fn:
        cnt[cur_line] += 1
tot = sum(cnt.values())
print("total", tot)
byfile = collections.Counter()
for (k, v) in cnt.items():
    byfile[k[0] if k else "?"] += v
print(dict(byfile))
for k, v in cnt.most_common(top):
    print(f"{k[0] if k else '?':18s}:{k[1] if k else 0:4d}  {v:5d}  {100*v/tot:5.1f}%")
