#!/bin/bash
for caps in "8,12,16,20,26" "8,11,14,18,24" "7,10,13,17,22" "6,9,12,16,22" "8,12,16,22,1000000" "9,13,17,21,27" "8,10,12,15,19" "10,14,18,22,28"; do
  echo "=== caps $caps"
  SQW_ADMM_CAPS=$caps python tools/gpu_admm_full.py cfg2 0 2>&1 | tail -1
done
