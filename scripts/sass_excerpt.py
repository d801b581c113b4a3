#!/usr/bin/env python
"""SASS evidence for profiles/: opcode histogram of one kernel and the instructions around the first occurrence of a marker
opcode (the hot loop). Usage: scripts/sass_excerpt.py <cubin> <kernel-substring> <marker-opcode> [lines-before] [lines-after]"""
import collections, re, subprocess, sys
cubin, kn, marker = sys.argv[1:4]
before = int(sys.argv[4]) if len(sys.argv) > 4 else 12
after = int(sys.argv[5]) if len(sys.argv) > 5 else 60
dis = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
fn = None; body = []
for ln in dis.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        fn = m.group(1); continue
    if fn and kn in fn and re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
        body.append(ln.rstrip())
ops = collections.Counter()
for ln in body:
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m: ops[m.group(1).split(".")[0]] += 1
print(f"# {kn}: {len(body)} SASS instructions (cuobjdump -sass, sm_100a)")
print("# opcode histogram: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(40)))
idx = next((i for i, ln in enumerate(body) if re.search(r"\b" + marker + r"\b", ln)), None)
if idx is not None:
    print(f"# around the first {marker}:")
    for ln in body[max(0, idx - before): idx + after]:
        print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln))
