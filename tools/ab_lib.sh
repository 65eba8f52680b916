#!/bin/bash
# Same-box A/B of two builds of libccvsq.so on the screen kernel alone: bash tools/ab_lib.sh <base.so> [workload] [reps]
BASE=$1; WL=${2:-c2}; REPS=${3:-2}
for i in $(seq $REPS); do
  CCVSQ_LIB=$BASE python tools/time_screen.py $WL 2>&1 | tail -1
  python tools/time_screen.py $WL 2>&1 | tail -1
done
