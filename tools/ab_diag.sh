#!/bin/bash
# A/B of several builds of libccvsq on the same box: per-entry-point timings (tools/diag_forward.py) and the bench
# value.  usage: tools/ab_diag.sh workload lib1.so lib2.so ...   ("-" = the in-tree default build)
WL=$1; shift
mkdir -p gpurun_out/ab
for rep in 1 2; do
  for lib in "$@"; do
    tag=$(basename $lib .so)
    if [ "$lib" == "-" ]; then unset CCVSQ_LIB; tag=default; else export CCVSQ_LIB=$PWD/$lib; fi
    python tools/diag_forward.py $WL > gpurun_out/ab/diag_${tag}_${WL}_$rep.txt 2>&1
    python bench.py --workload $WL --no-cpu-baseline > gpurun_out/ab/bench_${tag}_${WL}_$rep.json 2>/dev/null
    echo "== $tag rep $rep"; grep -E "composite full \(cached|stepwise assign|gather cm  |assign, no" gpurun_out/ab/diag_${tag}_${WL}_$rep.txt
    python -c "import json,sys; d=json.load(open('gpurun_out/ab/bench_${tag}_${WL}_$rep.json')); print('   bench value %.1f M/s  ms/step %.4f  assign %.4f  screen %.4f' % (d['value']/1e6, d['ms_per_step'], d['kernel_breakdown'].get('ccvsq_assign',{}).get('ms_per_step',0), d['kernel_breakdown'].get('ccvsq_screen',{}).get('ms_per_step',0)))"
  done
done
