#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r02_tma_ab.jsonl
timeout 600 python -m pytest tests/test_scan_gpu.py -x -q 2>&1 | tail -3
for mc in 0 8 2; do
 for dt in f32 bf16; do
  echo "== FM_SCAN_FWD16_TMA_MINCHUNKS=$mc $dt" >> gpurun_out/r02_tma_ab.jsonl
  FM_SCAN_FWD16_TMA_MINCHUNKS=$mc timeout 300 python tools/bench_vs_ref_cuda.py --dtype $dt --iters 30 --shapes configs1,stage0,stage1,long 2>&1 | cut -c1-110 >> gpurun_out/r02_tma_ab.jsonl
 done
done
cat gpurun_out/r02_tma_ab.jsonl
