#!/bin/bash
# round 2, visit a: baseline numbers on this pool + ncu of the HBM kernels at the 8x8 layouts + H2D probe
OUT=gpurun_out/r02a
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
python tools/time_screen.py c2 > $OUT/time_screen.txt 2>&1
python tools/time_screen.py c4 >> $OUT/time_screen.txt 2>&1
python tools/h2d_probe.py > $OUT/h2d_1rank.json 2> $OUT/h2d.err
python tools/h2d_probe.py --bind > $OUT/h2d_1rank_bound.json 2>> $OUT/h2d.err
for wl in c4 c3d512 c3; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cm4_kernel|rows4_kernel|rowsw_kernel" -s 6 -c 6 -o $OUT/stream_$wl -f \
      python tools/ncu_stream.py $wl > $OUT/ncu_stream_$wl.log 2>&1; echo "ncu stream $wl rc=$?"
done
python tools/diag_forward.py c4 > $OUT/diag_c4.txt 2>&1
python tools/diag_forward.py c3d512 > $OUT/diag_c3d512.txt 2>&1
python tools/diag_forward.py c3 > $OUT/diag_c3.txt 2>&1
tail -3 $OUT/time_screen.txt
