#!/bin/bash
O=gpurun_out/$1; mkdir -p $O
for i in 1 2; do
  python tools/time_screen.py c2 2>&1 | tail -1 | sed "s/^/new(preload): /" >> $O/time.txt
  CCVSQ_LIB=build/libccvsq_nopre.so python tools/time_screen.py c2 2>&1 | tail -1 | sed "s/^/structure only: /" >> $O/time.txt
  CCVSQ_LIB=build/libccvsq_prev.so python tools/time_screen.py c2 2>&1 | tail -1 | sed "s/^/prev: /" >> $O/time.txt
done
cat $O/time.txt
