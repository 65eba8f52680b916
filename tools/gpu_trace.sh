#!/bin/bash
# timeline + floor diagnostics of the screen kernel at c2: wide plan vs 64-column plan, with and without the candidate slow path
O=gpurun_out/$1; mkdir -p $O
python tools/trace_screen.py c2 > $O/trace_wide.txt 2>&1
CCVSQ_SCREEN_PLAN=64,3,2 python tools/trace_screen.py c2 > $O/trace_64.txt 2>&1
python tools/trace_screen.py c2 -1e28 > $O/trace_wide_noslow.txt 2>&1
for tau in 1.0 0.0 -1e28; do
  python tools/time_screen.py c2 $tau >> $O/time.txt 2>&1
  CCVSQ_SCREEN_PLAN=64,3,2 python tools/time_screen.py c2 $tau >> $O/time.txt 2>&1
done
cat $O/time.txt
tail -40 $O/trace_wide.txt
