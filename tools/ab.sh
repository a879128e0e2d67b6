#!/bin/bash
BIG=100000,100001,100002,100003,100004
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for lib in "$@"; do
  echo "=== $lib"
  SQW_LIB_PATH=$lib python tools/gpu_admm_full.py cfg2 0 2>&1 | tail -1
  SQW_LIB_PATH=$lib SQW_ADMM_CAPS=$BIG python tools/gpu_admm_full.py cfg2 0 2>&1 | tail -1
done
SQW_ADMM_PROFILE=1 python tools/gpu_admm_full.py cfg2 0 2>&1 | grep 'sqw profile'
