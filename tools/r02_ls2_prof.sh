#!/bin/bash
# ncu --set full of the pipelined lane-serial backward from the stand-alone harness (after the plain run exited 0)
mkdir -p gpurun_out
FM_SCAN_BWD_LS2_NW=${NW:-1} timeout 120 tools/ubench/ls2_check 8 4096 | tail -1 || exit 1
FM_SCAN_BWD_LS2_NW=${NW:-1} ncu --set full --clock-control none --import-source on -k regex:scan_bwd_ls2 -s 1 -c 1 -o gpurun_out/r02_ls2 -f tools/ubench/ls2_check 8 4096 > gpurun_out/r02_ls2_ncu.log 2>&1
tail -2 gpurun_out/r02_ls2_ncu.log
ls -la gpurun_out/r02_ls2.ncu-rep
