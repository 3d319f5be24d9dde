#!/bin/bash
# scripts/ab_run.sh OUTFILE lib1.so lib2.so ...   ("default" = the in-tree library): one scripts/ab_measure.py line per build
OUT=$1; shift
for lib in "$@"; do
  if [ "$lib" = "default" ]; then timeout 300 python scripts/ab_measure.py >> $OUT 2>&1
  else RFWB200_LIB=$lib timeout 300 python scripts/ab_measure.py >> $OUT 2>&1; fi
done
cat $OUT
