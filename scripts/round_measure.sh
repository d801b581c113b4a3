#!/bin/bash
# Round-end measurement pass on the GPU box: parity tests, bench (both arms), launch list, ncu captures, config benches.
tag=${1:-r2}
out=gpurun_out
timeout 400 python -m pytest tests -m gpu -q > $out/pytest_gpu_$tag.log 2>&1; tail -2 $out/pytest_gpu_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; tail -c 200 $out/bench_$tag.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_${tag}_reference.json 2>> $out/bench_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --unique 32 --no-cpu --latency-pairs 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_cov_leaf_kernel -s 1 -c 1 -o $out/prof_knn_$tag -f python bench.py --steps 2 --warmup 1 --unique 32 --no-cpu --latency-pairs 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:align_kernel -s 1 -c 1 -o $out/prof_align_$tag -f python bench.py --steps 2 --warmup 1 --unique 32 --no-cpu --latency-pairs 4 > /dev/null 2>&1
python scripts/bench_configs.py ${CONFIGS:-c1 c3 c4 fit pre submap c5} > $out/configs_$tag.jsonl 2> $out/configs_$tag.err; tail -c 300 $out/configs_$tag.err
ls -la $out | tail -12
