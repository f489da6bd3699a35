#!/bin/bash
# tests + bench lines + knn timing (one gpurun call)
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "=== bench dist"
timeout 600 python bench.py --workload dist --steps 5 --warmup 3 > gpurun_out/bench_j_dist.json 2> gpurun_out/bench_j_dist.err; cut -c1-1500 gpurun_out/bench_j_dist.json
echo "=== knn timing"
timeout 600 python scripts/knn_time.py 2>&1 | tail -5
