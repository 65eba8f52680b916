#!/bin/bash
# A/B of screen plans (CCVSQ_SCREEN_PLAN=bn,nacc,abuf) on the same box.  usage: tools/ab_plan.sh workload plan...
WL=$1; shift
mkdir -p gpurun_out/ab
for rep in 1 2; do
  for plan in default "$@"; do
    if [ "$plan" == "default" ]; then unset CCVSQ_SCREEN_PLAN; else export CCVSQ_SCREEN_PLAN=$plan; fi
    python bench.py --workload $WL --no-cpu-baseline > gpurun_out/ab/plan_${plan}_${WL}_$rep.json 2>/dev/null
    python -c "
import json; d=json.load(open('gpurun_out/ab/plan_${plan}_${WL}_$rep.json')); kb=d['kernel_breakdown']
print('plan %-8s $WL rep $rep: value %.1f M/s  ms/step %.4f  screen %.4f ms (%.1f TF/s, %.1f%%)' % ('$plan', d['value']/1e6, d['ms_per_step'], kb['ccvsq_screen']['ms_per_step'], d['roofline']['achieved'], 100*d['roofline']['frac']))"
  done
done
