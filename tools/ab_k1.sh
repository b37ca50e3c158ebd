#!/bin/bash
# A/B of K1 builds: tools/ab_k1.sh <n_rep> <kinds> lib1.so lib2.so ...   (development aid)
nrep=$1; kinds=$2; shift 2
for lib in "$@"; do
  cp "$lib" literate_b200/_lib/libliterate_b200.so
  echo "== $lib"
  timeout 300 python tools/k1_bench.py $nrep 0 $kinds 10 2>&1 | grep -E "GBps|rror"
done
