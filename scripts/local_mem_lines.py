#!/usr/bin/env python
"""Where a kernel touches local memory (spills, demoted arrays): STL/LDL instructions per source line.
Usage: scripts/local_mem_lines.py <cubin> <kernel-substring>"""
import re, subprocess, sys
cubin, kn = sys.argv[1:3]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
cur = None; fn = None; res = {}; n = 0
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m: fn = m.group(1); continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if fn and kn in fn:
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln): n += 1
        if re.search(r'\b(STL|LDL)\b', ln): res.setdefault(cur, []).append(ln.strip()[:90])
print(kn, 'instructions', n)
for k, v in sorted(res.items(), key=lambda kv: (kv[0] or ('', 0))):
    print(k, len(v), v[0])
