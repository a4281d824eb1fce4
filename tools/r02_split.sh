#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scan_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/r02_split_tests.log; tail -6 gpurun_out/r02_split_tests.log
for dt in f32 bf16; do python tools/bench_vs_ref_cuda.py --shapes long,configs1 --dtype $dt --iters 10; done 2>&1 | cut -c1-330 | tee gpurun_out/r02_split_bench.jsonl
for j in 4 8 12 16; do FM_SCAN_FWD16_NSEG=$j python tools/bench_vs_ref_cuda.py --shapes long --dtype f32 --iters 10 | cut -c1-120; done 2>&1 | tee -a gpurun_out/r02_split_bench.jsonl
