#!/bin/bash
# A/B builds of the library on cfg-2 on one box: usage tools/ab.sh lib1.so lib2.so ...
for rep in 1 2; do
for lib in "$@"; do
  echo "=== $lib (rep $rep)"
  SQW_LIB_PATH=$lib python tools/gpu_admm_full.py cfg2 0 2>&1 | tail -2
done
done
