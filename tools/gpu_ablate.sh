#!/bin/bash
O=gpurun_out/$1; mkdir -p $O
for a in 0 1 2 3 4 7; do
  CCVSQ_SCREEN_ABLATE=$a python tools/time_screen.py c2 2>&1 | tail -1 | sed "s/^/ablate=$a wide: /" >> $O/ablate.txt
  CCVSQ_SCREEN_ABLATE=$a CCVSQ_SCREEN_PLAN=64,3,2 python tools/time_screen.py c2 2>&1 | tail -1 | sed "s/^/ablate=$a n64:  /" >> $O/ablate.txt
done
cat $O/ablate.txt
