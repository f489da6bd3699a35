#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== cgroup / cpu"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; nproc; python -c "import os; print('affinity', len(os.sched_getaffinity(0)))"; lscpu | grep -E "Socket|Core|Thread|NUMA node\(s\)"
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "=== ncu full: dist_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^dist_kernel -s 1 -c 1 -f -o gpurun_out/prof_dist2 \
    python bench.py --workload dist --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_dist2.log 2>&1
tail -2 gpurun_out/ncu_dist2.log | cut -c1-200
echo "=== omp threads test"
for t in 16 32 64 128; do OMP_NUM_THREADS=$t timeout 300 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print($t, d['value'], d['cpu_baseline']['sample'][:60])"; done
