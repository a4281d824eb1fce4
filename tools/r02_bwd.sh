#!/bin/bash
# round-2: lane-serial backward -- parity suite, then A/B timing against the row-pair kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scan_gpu.py tests/test_ss2d_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/r02_bwd_tests.log
tail -4 gpurun_out/r02_bwd_tests.log
rm -f gpurun_out/r02_bwd_ab.jsonl
for cfg in "FM_SCAN_BWD_LS=0" "FM_SCAN_BWD_LS_NW=1" "FM_SCAN_BWD_LS_NW=2" "FM_SCAN_BWD_LS_NW=4"; do
  for dt in f32 bf16; do
    echo "== $cfg $dt" >> gpurun_out/r02_bwd_ab.jsonl
    env $cfg timeout 300 python tools/bench_vs_ref_cuda.py --dtype $dt --iters 20 >> gpurun_out/r02_bwd_ab.jsonl 2>&1
  done
done
cat gpurun_out/r02_bwd_ab.jsonl | cut -c1-400
