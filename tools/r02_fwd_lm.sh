#!/bin/bash
# round-2: forward lane-map experiment (quarter warps share B/C packets) -- parity with the map on, then A/B
mkdir -p gpurun_out
FM_SCAN_FWD16_LM=1 timeout 900 python -m pytest tests/test_scan_gpu.py -x -q 2>&1 | tail -3
FM_SCAN_FWD16_LM=1 FM_SCAN_FWD16_SPL=4 timeout 900 python -m pytest tests/test_scan_gpu.py -x -q -k "golden or fixtures or configs1 or every_length" 2>&1 | tail -3
python tools/ab.py --which fwd --shapes configs1,stage0,stage1,stage2,train_s0 --dtypes f32,bf16 --iters 20 \
  --env "" --env "FM_SCAN_FWD16_LM=1" --env "FM_SCAN_FWD16_SPL=4" --env "FM_SCAN_FWD16_SPL=4,FM_SCAN_FWD16_LM=1" \
  --env "FM_SCAN_FWD16_SPL=2,FM_SCAN_FWD16_LM=1" > gpurun_out/r02_fwd_lm_ab.jsonl 2>&1
cat gpurun_out/r02_fwd_lm_ab.jsonl
