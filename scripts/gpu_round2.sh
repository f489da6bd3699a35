#!/bin/bash
# Round-end GPU pass: tests, smoke, bench (+ reference arm), ncu launch list, full captures of the kernels added this round,
# compute-sanitizer on the new paths.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
echo "=== nvidia-smi"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
echo "=== pytest -m gpu"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -4 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
echo "=== bench reference"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
echo "=== ncu launch list"
KREGEX='regex:dist_kernel|dist_jmle_kernel|sketch_kernel|planes_kernel|card_kernel|range_kernel|pack_kernel|mark_starts_kernel|cardinality_kernel|fa_|union_kernel|compress_kernel|knn_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 1200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200; wc -l gpurun_out/launches.csv
echo "=== ncu full: dist_kernel, sketch_kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^dist_kernel -s 1 -c 1 -f -o gpurun_out/prof_dist \
    python bench.py --workload dist --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_dist.log 2>&1; tail -1 gpurun_out/ncu_dist.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_kernel -s 1 -c 1 -f -o gpurun_out/prof_sketch \
    python bench.py --workload sketch --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_sketch.log 2>&1; tail -1 gpurun_out/ncu_sketch.log | cut -c1-200
echo "=== ncu full: fasta kernels + union"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_emit_kernel -s 3 -c 1 -f -o gpurun_out/prof_fa_emit python scripts/setops_run.py fasta > gpurun_out/ncu_fa.log 2>&1; tail -1 gpurun_out/ncu_fa.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_summary_kernel -s 3 -c 1 -f -o gpurun_out/prof_fa_summary python scripts/setops_run.py fasta > gpurun_out/ncu_fa2.log 2>&1; tail -1 gpurun_out/ncu_fa2.log | cut -c1-200
ls -la gpurun_out
