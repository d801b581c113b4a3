import sys; sys.path.insert(0,'/root/repo')
from riv_slam_b200 import fast_apdgicp as F
H=F.Handle(0)
for n in (200000, 1000000, 8000000):
    r=H.bench_streaming(n,5)
    print(n, {k:(round(g),round(g/6545.6,2),round(m*1e3,1)) for k,(g,m) in r.items()})
