#!/usr/bin/env python
"""C2 step (1000 pairs x 5000 points) through apd_odometry_align with the scans already in device memory (option
pipeline_device_input): the two-stream chunk pipeline without the host-to-device copy, against the single-stream
create / prepare / align_pairs sequence bench.py times as `value`."""
import ctypes as C, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from riv_slam_b200 import fast_apdgicp as F
from bench import LAUNCH_PARAMS, make_workload, to_pointxyzi, N_POINTS

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
opts = [a.split("=") for a in sys.argv[3:]]
scans, order = make_workload(P, 96, 0, 16)
pts, off = to_pointxyzi(scans, order)
host = torch.from_numpy(pts.view(np.uint8).reshape(-1)).pin_memory()
dev = host.cuda()
H = F.Handle(0); H.set_params(**LAUNCH_PARAMS)
for k, v in opts:
    H.set_option(k, float(v))
L = H.L
ip = C.POINTER(C.c_int32)
res = np.zeros(P, dtype=F.RESULT_DTYPE)
res_pin = torch.from_numpy(res.view(np.uint8).reshape(-1)).pin_memory()
def run(ptr, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        H.check(L.apd_odometry_align(H.h, C.c_void_p(ptr), off.ctypes.data_as(ip), P + 1, 32, None, C.c_void_p(res_pin.data_ptr())))
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, ptr, flag in (("host input (e2e)", host.data_ptr(), 0), ("device input, pipelined", dev.data_ptr(), 1)):
    H.set_option("pipeline_device_input", flag)
    run(ptr, 3)
    ms = run(ptr, steps)
    view = np.frombuffer(memoryview(res_pin.numpy()), dtype=F.RESULT_DTYPE)
    print(f"{name:28s} {ms:7.3f} ms per {P} pairs = {P / ms:7.1f} k reg/s   converged {float(np.mean(view['converged'] != 0)):.3f}")
