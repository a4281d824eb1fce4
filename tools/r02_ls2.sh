#!/bin/bash
# round-2: pipelined lane-serial backward (fm_scan_bwd_ls2.cuh) -- stand-alone check against the first lane-serial kernel + timing
mkdir -p gpurun_out
out=gpurun_out/r02_ls2_check.log
: > $out
bin=tools/ubench/ls2_check
for nw in 1 2; do
  echo "== NW=$nw" >> $out
  FM_SCAN_BWD_LS2_NW=$nw timeout 120 $bin 8 4096 >> $out 2>&1
done
for a in "4 4096" "12 4096" "32 1024" "32 256" "8 4092" "8 1000" "32 64" "32 16" "2 20480"; do timeout 120 $bin $a | tail -1 >> $out 2>&1; done
cat $out
