#!/bin/bash
# One GPU visit: parity tests, benches, ncu launch list and full captures of the screen kernel.
# Usage (from the repo root, on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --workload c4 --no-cpu-baseline > $OUT/bench_c4.json 2> $OUT/bench_c4.err; echo "bench c4 rc=$?"
timeout 600 python bench.py --workload c3 --no-cpu-baseline > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "bench c3 rc=$?"
timeout 600 python bench.py --workload train --no-cpu-baseline > $OUT/bench_train.json 2> $OUT/bench_train.err; echo "bench train rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_c2.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:screen_kernel -s 3 -c 1 -o $OUT/screen_c4 -f \
    python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_c4.log 2>&1; echo "ncu full c4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:screen_kernel -s 3 -c 1 -o $OUT/screen_c2 -f \
    python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_c2full.log 2>&1; echo "ncu full c2 rc=$?"
python tools/show_bench.py $OUT/bench_c2.json $OUT/bench_c4.json $OUT/bench_c3.json $OUT/bench_train.json 2>&1 | tail -60
