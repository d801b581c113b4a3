#!/bin/bash
# usage: sweep_option.sh <option-name> v1 v2 ...   (bench.py --option name=value per run)
name=$1; shift
for v in "$@"; do
  python bench.py --no-cpu --unique 48 --latency-pairs 40 --option $name=$v 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name=$v', round(d['value']), round(d['e2e']['value']), round(d['p50_align_latency_ms'],3), {k:round(v,2) for k,v in d['phases_ms_per_step'].items()})"
done
