#!/bin/bash
O=gpurun_out/$1; mkdir -p $O
for a in 0 1 32 64 96 3; do
  CCVSQ_SCREEN_ABLATE=$a python tools/time_screen.py c2 2>&1 | tail -1 | sed "s/^/ablate=$a wide: /" >> $O/ablate.txt
done
cat $O/ablate.txt
