import sys, time, json, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from riv_slam_b200 import fast_apdgicp as F
from test_preprocess import raw_scan, oracle_pipeline
H = F.Handle(0)
for nrep, label in ((3, 10202), (12, 41085)):
    raw = np.concatenate([raw_scan(60 + i, 2000) for i in range(nrep)])
    for method in ("VOXELGRID", "APPROX_VOXELGRID"):
        out = F.preprocess(H, raw, downsample_method=method)
        ts = []
        for _ in range(30):
            t0 = time.perf_counter(); out = F.preprocess(H, raw, downsample_method=method); ts.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); want = oracle_pipeline(raw, downsample_method=method); tc = time.perf_counter() - t0
        print(json.dumps({"config": f"8(f)-4 preprocessing filters (distance, {method} 0.1 m, RadiusOutlierRemoval 0.8 m / 2)", "points_in": len(raw), "points_out": len(out),
                          "gpu_ms_p50": round(float(np.median(ts)) * 1e3, 4), "cpu_oracle_ms": round(tc * 1e3, 1), "bit_exact": out.tobytes() == want.tobytes()}))
