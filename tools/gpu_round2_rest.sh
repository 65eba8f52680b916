#!/bin/bash
# The K = 16384 workloads of the round-2 bundle (bench lines + the full ncu capture of the screen kernel at c4).
TAG=${1:-r05b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for wl in c4 c3 c3d512; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-extras > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "bench $wl rc=$?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:screen_kernel -s 3 -c 1 -o $OUT/screen_c4 -f \
    python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_c4.log 2>&1; echo "ncu full c4 rc=$?"
python tools/show_bench.py $OUT/bench_c4.json $OUT/bench_c3.json $OUT/bench_c3d512.json 2>&1 | tail -12
