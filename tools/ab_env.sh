#!/bin/bash
# A/B of one build under two environments (e.g. CCVSQ_NO_FUSE=1 vs default): bench value and kernel breakdown.
# usage: tools/ab_env.sh "VAR=1" workload...
VAR=$1; shift
mkdir -p gpurun_out/ab
for rep in 1 2; do
  for wl in "$@"; do
    python bench.py --workload $wl --no-cpu-baseline > gpurun_out/ab/env_new_${wl}_$rep.json 2>gpurun_out/ab/env_new_${wl}_$rep.err
    env $VAR python bench.py --workload $wl --no-cpu-baseline > gpurun_out/ab/env_old_${wl}_$rep.json 2>/dev/null
    for t in new old; do
      python -c "
import json; d=json.load(open('gpurun_out/ab/env_${t}_${wl}_$rep.json')); kb=d['kernel_breakdown']
print('$t $wl rep $rep: value %.1f M/s  ms/step %.4f  screen %.4f ms (%.1f TF/s)  e2e %.1f M/s' % (d['value']/1e6, d['ms_per_step'], kb.get('ccvsq_screen',{}).get('ms_per_step',0), d['roofline']['achieved'] or 0, d['e2e']['value']/1e6))"
    done
  done
done
