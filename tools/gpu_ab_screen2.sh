#!/bin/bash
# same-box A/B of two builds of the library on the screen kernel (c2), then the GPU tests of the current build
O=gpurun_out/$1; mkdir -p $O
for i in 1 2; do
  python tools/time_screen.py c2 2>&1 | tail -1 | sed "s/^/new: /" >> $O/time.txt
  [ -f build/libccvsq_prev.so ] && CCVSQ_LIB=build/libccvsq_prev.so python tools/time_screen.py c2 2>&1 | tail -1 | sed "s/^/prev: /" >> $O/time.txt
done
python tools/time_screen.py c1 2>&1 | tail -1 | sed "s/^/new c1: /" >> $O/time.txt
cat $O/time.txt
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
