#!/bin/bash
# One GPU-box round: tests, smoke, bench, ncu launch list + full captures.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
KREGEX='regex:dist_kernel|dist_jmle_kernel|sketch_kernel|planes_kernel|card_kernel|range_kernel|pack_kernel|mark_starts_kernel|cardinality_kernel'
echo "=== nvidia-smi"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
echo "=== host"; nproc; grep -m1 "model name" /proc/cpuinfo
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
if [ "${SKIP_BENCH:-0}" != "1" ]; then
echo "=== bench"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
echo "=== bench reference"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log | cut -c1-300
echo "=== ncu full: dist_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^dist_kernel -s 1 -c 1 -f -o gpurun_out/prof_dist \
    python bench.py --workload dist --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_dist.log 2>&1
tail -2 gpurun_out/ncu_dist.log | cut -c1-200
echo "=== ncu full: sketch_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sketch_kernel -s 1 -c 1 -f -o gpurun_out/prof_sketch \
    python bench.py --workload sketch --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_sketch.log 2>&1
tail -2 gpurun_out/ncu_sketch.log | cut -c1-200
fi
ls -la gpurun_out
