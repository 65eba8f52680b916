#!/bin/bash
O=gpurun_out/$1; mkdir -p $O
for a in 0 1 32 3 131 11; do
  CCVSQ_SCREEN_ABLATE=$a python tools/time_screen.py c2 2>&1 | tail -1 | sed "s/^/ablate=$a wide: /" >> $O/ablate.txt
done
cat $O/ablate.txt
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
