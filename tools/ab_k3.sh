#!/bin/bash
# A/B of K3 builds: tools/ab_k3.sh <chains> <iters> <variants> lib1.so lib2.so ...   (development aid)
chains=$1; iters=$2; variants=$3; shift 3
for lib in "$@"; do
  cp "$lib" literate_b200/_lib/libliterate_b200.so
  echo "== $lib"
  timeout 300 python tools/k3_bench.py $chains $iters 1000 $variants 2>&1 | grep -E "variant|K3PROF|Error|error"
done
