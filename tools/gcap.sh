#!/bin/bash
for cap in 1000 72 48 36 24; do
  echo "=== gcap $cap"
  SQW_ADMM_GCAP=$cap python tools/gpu_admm_full.py cfg2 0 2>&1 | tail -1
done
