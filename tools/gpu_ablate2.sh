#!/bin/bash
O=gpurun_out/$1; mkdir -p $O
for a in 3 11 19 27; do
  CCVSQ_SCREEN_ABLATE=$a python tools/time_screen.py c2 2>&1 | tail -1 | sed "s/^/ablate=$a wide: /" >> $O/ablate.txt
done
cat $O/ablate.txt
CCVSQ_SCREEN_ABLATE=11 python tools/trace_screen.py c2 > $O/trace_floor_nost.txt 2>&1; grep "leader CTAs\|loaders,\|epilogue," $O/trace_floor_nost.txt; grep -A9 "CTA 0, sweep 3" $O/trace_floor_nost.txt
CCVSQ_SCREEN_ABLATE=19 python tools/trace_screen.py c2 > $O/trace_floor_nold.txt 2>&1; grep "leader CTAs\|loaders,\|epilogue," $O/trace_floor_nold.txt; grep -A9 "CTA 0, sweep 3" $O/trace_floor_nold.txt
