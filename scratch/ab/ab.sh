#!/bin/bash
# A/B of two builds of the library on the same box: scratch/ab/lib_head.so vs scratch/ab/lib_new.so (measurement script)
for rep in 1 2 3; do
  for v in head new; do
    cp scratch/ab/lib_$v.so curious_b200/libcurious_b200.so
    echo -n "$v: "; CUR_ROWS_PAIR=${PAIR:-0} B=${B:-256} timeout 120 python scratch/ddpg_bench.py 2>&1 | tail -1
  done
done
cp scratch/ab/lib_new.so curious_b200/libcurious_b200.so
