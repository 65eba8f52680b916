#!/bin/bash
# same-box timing of the screen kernel: current build (wide plan / 64 plan), ablations, c4 both plans; then the GPU tests
O=gpurun_out/$1; mkdir -p $O
for a in 0 1 3; do
  CCVSQ_SCREEN_ABLATE=$a python tools/time_screen.py c2 2>&1 | tail -1 | sed "s/^/ablate=$a wide: /" >> $O/time.txt
done
CCVSQ_SCREEN_PLAN=64,3,2 python tools/time_screen.py c2 2>&1 | tail -1 | sed "s/^/n64: /" >> $O/time.txt
python tools/time_screen.py c4 2>&1 | tail -1 | sed "s/^/c4 default: /" >> $O/time.txt
CCVSQ_SCREEN_BN=128 python tools/time_screen.py c4 2>&1 | tail -1 | sed "s/^/c4 wide: /" >> $O/time.txt
python tools/time_screen.py c3d512 2>&1 | tail -1 | sed "s/^/c3d512 default: /" >> $O/time.txt
CCVSQ_SCREEN_BN=128 python tools/time_screen.py c3d512 2>&1 | tail -1 | sed "s/^/c3d512 wide: /" >> $O/time.txt
cat $O/time.txt
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
python tools/tau_cost.py c2 > $O/tau_c2.txt 2>&1; cat $O/tau_c2.txt
