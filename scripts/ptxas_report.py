#!/usr/bin/env python
"""Per-kernel register / spill / shared-memory report from `nvcc -Xptxas -v` (python scripts/ptxas_report.py [filter])."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from riv_slam_b200 import build as B
def main():
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    cmd = [B._nvcc()] + B.NVCC_FLAGS + ["-Xptxas", "-v", "-o", "/dev/null"] + [os.path.join(B.CSRC, s) for s in B.SOURCES]
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    err = subprocess.run(cmd, capture_output=True, text=True).stderr
    cur = None
    for line in err.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::|apd::|void ", "", cur).split("(")[0]
            continue
        if cur and "spill" in line:
            spill = line.strip()
        m = re.search(r"Used (\d+) registers.*", line)
        if m and cur:
            if flt in cur:
                print(f"{cur:70s} regs {m.group(1):>3s}  {spill}  {line.split('barriers,')[-1].strip() if 'smem' in line else ''}")
            cur = None
if __name__ == "__main__":
    main()
