#!/bin/bash
# Quick GPU iteration: parity subset + c4/c2 bench + one full ncu capture of the screen kernel at c4.
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -2 $OUT/pytest_gpu.log
timeout 300 python bench.py --workload c4 --steps 10 --no-cpu-baseline > $OUT/bench_c4.json 2> $OUT/bench_c4.err; echo "bench c4 rc=$?"
timeout 300 python bench.py --workload c2 --steps 20 --no-cpu-baseline > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench c2 rc=$?"
if [ "$2" != "noncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:screen_kernel -s 3 -c 1 -o $OUT/screen_c4 -f \
    python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_c4.log 2>&1; echo "ncu full c4 rc=$?"
fi
python tools/show_bench.py $OUT/bench_c4.json $OUT/bench_c2.json 2>&1 | tail -20
