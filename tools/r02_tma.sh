#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scan_gpu.py tests/test_ss2d_gpu.py -x -q 2>&1 | tail -8 > gpurun_out/r02_tma_tests.log; tail -4 gpurun_out/r02_tma_tests.log
for dt in f32 bf16; do timeout 300 python tools/bench_vs_ref_cuda.py --dtype $dt --iters 20; done 2>&1 | cut -c1-200 | tee gpurun_out/r02_tma_bench.jsonl
