#!/bin/bash
# same-box A/B of two library builds on the HER kernel alone (scratch/her_bench.py); see ab.sh
for rep in 1 2 3; do
  for v in head new; do
    cp scratch/ab/lib_$v.so curious_b200/libcurious_b200.so
    echo -n "$v: "; timeout 200 python scratch/her_bench.py 2>&1 | tail -2 | tr '\n' ' '; echo
  done
done
cp scratch/ab/lib_new.so curious_b200/libcurious_b200.so
