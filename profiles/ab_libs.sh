#!/bin/bash
# Same-box A/B of library variants: per-kernel-class CUDA-event times of a full 50-step decode (batch 64, graphs off).
#   profiles/ab_libs.sh <precision> <name>=<lib.so> ...      ("cur" = the in-tree library)
prec=$1; shift
for kv in "$@"; do
  name=${kv%%=*}; lib=${kv#*=}
  if [ "$lib" = "cur" ]; then unset SELFTOK_B200_LIB; else export SELFTOK_B200_LIB=$PWD/$lib; fi
  echo "== $name"
  timeout 300 python profiles/step_classes.py $prec 50 2>&1 | tail -1
done
