import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import LAUNCH_PARAMS
from riv_slam_b200 import datagen, fast_apdgicp as F
n_pairs, uniq = 4096, 16
base = [datagen.make_pair(4, i, n_src=5000) for i in range(uniq)]
srcs = [base[i % uniq][0] for i in range(n_pairs)]
ps, os_ = F._ragged(srcs)
for opt in ([], [("smem_build", 0)]):
    H = F.Handle(0)
    H.set_params(**LAUNCH_PARAMS)
    for k, v in opt:
        H.set_option(k, v)
    for rep in range(3):
        t0 = time.perf_counter(); S = F.CloudSet(H, (ps, os_)); H.synchronize(); t1 = time.perf_counter()
        S.prepare(); H.synchronize(); t2 = time.perf_counter()
        l0 = H.launch_count()
        S.destroy()
        print(opt, "create %.1f ms prepare %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
