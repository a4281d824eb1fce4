#!/bin/bash
# parity suite under the tuning knobs of the new kernels (every variant must give the oracle's results)
mkdir -p gpurun_out
for env in "FM_SCAN_BWD_LS2_NW=2" "FM_SCAN_BWD_LS2_NW=4" "FM_SCAN_FWD16_LM=0" "FM_SCAN_FWD16_LM=1 FM_SCAN_FWD16_SPL=4" "FM_SCAN_BWD_LS_MINUNITS=0 FM_SCAN_BWD_LS2_MINL=0" "FM_SCAN_FWD16_SPL=2 FM_SCAN_FWD16_KT=1"; do
  echo "== $env"
  env $env timeout 900 python -m pytest tests/test_scan_gpu.py tests/test_ss2d_gpu.py -x -q 2>&1 | tail -2
done 2>&1 | tee gpurun_out/r02_variants.log
