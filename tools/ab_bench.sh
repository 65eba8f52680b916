#!/bin/bash
# A/B of two builds of libccvsq on the same box: alternating bench runs.  usage: tools/ab_bench.sh other.so workload...
OTHER=$1; shift
mkdir -p gpurun_out/ab
for rep in 1 2; do
  for wl in "$@"; do
    python bench.py --workload $wl --no-cpu-baseline > gpurun_out/ab/new_${wl}_$rep.json 2>/dev/null
    CCVSQ_LIB=$OTHER python bench.py --workload $wl --no-cpu-baseline > gpurun_out/ab/old_${wl}_$rep.json 2>/dev/null
  done
done
python tools/show_bench.py -v gpurun_out/ab/*.json
