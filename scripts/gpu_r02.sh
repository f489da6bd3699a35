#!/bin/bash
# One GPU-box round (round 2): selected stages via $STAGES (space separated): tests smoke bench ref emu4 emu5 ncu_list ncu_dist ncu_planes ncu_jmle
set -u
mkdir -p gpurun_out
STAGES="${STAGES:-tests smoke bench ref}"
TAG="${TAG:-r02}"
has() { [[ " $STAGES " == *" $1 "* ]]; }
echo "=== nvidia-smi"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
echo "=== host"; nproc; grep -m1 "model name" /proc/cpuinfo; cat /sys/fs/cgroup/cpu.max 2>/dev/null; free -g | head -2
if has tests; then
echo "=== pytest -m gpu"
timeout 1800 python -m pytest tests -x -q -m gpu ${PYTEST_ARGS:-} 2>&1 | tail -25
fi
if has smoke; then
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
fi
if has bench; then
echo "=== bench"
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -12 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
fi
if has ref; then
echo "=== bench reference"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; tail -3 gpurun_out/${TAG}_bench_ref.err; cat gpurun_out/${TAG}_bench_ref.json
fi
if has emu4; then
echo "=== emulated rank of C4"
timeout 900 python bench.py --emulate-world 8 --emulate-rank ${EMU_RANK:-0} --only c4 > gpurun_out/${TAG}_c4_emu.json 2> gpurun_out/${TAG}_c4_emu.err; tail -8 gpurun_out/${TAG}_c4_emu.err; cat gpurun_out/${TAG}_c4_emu.json
fi
if has emu5; then
echo "=== emulated rank of C5"
timeout 1200 python bench.py --emulate-world 8 --emulate-rank ${EMU_RANK:-0} --only c5 > gpurun_out/${TAG}_c5_emu.json 2> gpurun_out/${TAG}_c5_emu.err; tail -8 gpurun_out/${TAG}_c5_emu.err; cat gpurun_out/${TAG}_c5_emu.json
fi
KREGEX='regex:dist_kernel|dist_jmle_kernel|dist_sweep|dist_estim|sketch_kernel|planes_kernel|card_kernel|range_kernel|pack_kernel|mark_starts_kernel|cardinality_kernel'
if has ncu_list; then
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_bench.log | cut -c1-300
fi
if has ncu_dist; then
echo "=== ncu full: dist kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^dist_|dist_sweep|dist_estim" -s ${NCU_SKIP:-2} -c ${NCU_COUNT:-2} -f -o gpurun_out/${TAG}_prof_dist \
    python bench.py --workload dist --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu_dist.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_dist.log | cut -c1-200
fi
if has ncu_planes; then
echo "=== ncu full: planes_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:planes_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_planes \
    python bench.py --workload dist --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu_planes.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_planes.log | cut -c1-200
fi
if has ncu_jmle; then
echo "=== ncu full: joint-MLE kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:jmle" -s 1 -c ${NCU_COUNT:-2} -f -o gpurun_out/${TAG}_prof_jmle \
    python scripts/jmle_run.py > gpurun_out/${TAG}_ncu_jmle.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_jmle.log | cut -c1-200
fi
if has ncu_sketch; then
echo "=== ncu full: sketch_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sketch_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_sketch \
    python bench.py --workload sketch --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu_sketch.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_sketch.log | cut -c1-200
fi
ls -la gpurun_out | tail -20
