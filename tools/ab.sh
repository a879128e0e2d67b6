#!/bin/bash
# A/B builds of the library on cfg-2: plain timing, profile in phased mode, then single-phase
# (constant G) timing and profile.   usage: tools/ab.sh lib1.so lib2.so ...
BIG=100000,100001,100002,100003,100004
for rep in 1 2; do
for lib in "$@"; do
  echo "=== $lib (rep $rep)"
  SQW_LIB_PATH=$lib python tools/gpu_admm_full.py cfg2 0 2>&1 | tail -1
  SQW_LIB_PATH=$lib SQW_ADMM_CAPS=$BIG python tools/gpu_admm_full.py cfg2 0 2>&1 | tail -1
done
done
for lib in "$@"; do
  echo "=== $lib profiles: phased, single-phase"
  SQW_LIB_PATH=$lib SQW_ADMM_PROFILE=1 python tools/gpu_admm_full.py cfg2 0 2>&1 | grep -A1 "cta 0:"
  SQW_LIB_PATH=$lib SQW_ADMM_CAPS=$BIG SQW_ADMM_PROFILE=1 python tools/gpu_admm_full.py cfg2 0 2>&1 | grep -A1 "cta 0:"
done
