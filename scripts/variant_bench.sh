#!/bin/bash
# runs bench.py once per library variant under variants/ (kernel-shape experiments); prints value / phases per variant
for lib in "" variants/*.so; do
  if [ -n "$lib" ]; then export APDGICP_B200_LIB=$PWD/$lib; else unset APDGICP_B200_LIB; fi
  python bench.py --no-cpu --unique 48 --latency-pairs 40 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('${lib:-default}', round(d['value']), round(d['e2e']['value']), round(d['p50_align_latency_ms'],3), {k:round(v,2) for k,v in d['phases_ms_per_step'].items()})"
done
