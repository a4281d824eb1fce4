#!/bin/bash
# whole GPU suite + the default bench line, as the driver runs them at round end
mkdir -p gpurun_out; rm -f gpurun_out/model_parity.jsonl
( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | grep -v Warning | tail -12 > gpurun_out/r02_all_tests.log; tail -8 gpurun_out/r02_all_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2
python tools/bench_vs_ref_cuda.py --shapes long --dtype f32,bf16 --iters 10 2>&1 | cut -c1-330
bash tools/r02_bench.sh 1
