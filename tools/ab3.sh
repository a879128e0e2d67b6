#!/bin/bash
BIG=100000,100001,100002,100003,100004
for lib in "$@"; do
  echo "=== $lib: single-phase, phased"
  SQW_LIB_PATH=$lib SQW_ADMM_CAPS=$BIG SQW_ADMM_PROFILE=1 python tools/gpu_admm_full.py cfg2 0 2>&1 | grep -A2 "cta 0:"
  SQW_LIB_PATH=$lib SQW_ADMM_PROFILE=1 python tools/gpu_admm_full.py cfg2 0 2>&1 | grep -A2 "cta 0:"
done
