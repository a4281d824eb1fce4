#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v Warning | tail -6
python tools/model_bench.py train --steps 3 --warmup 2 > gpurun_out/r02_train_ln.json 2> gpurun_out/r02_train_ln.err; tail -c 1600 gpurun_out/r02_train_ln.json
python tools/model_bench.py long --steps 5 --warmup 2 > gpurun_out/r02_long3.json 2> gpurun_out/r02_long3.err; cat gpurun_out/r02_long3.json
