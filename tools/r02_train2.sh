#!/bin/bash
# round-2: fused conv/unfold backward on the training path -- parity (SS2D + model level), then the configs[3] training step per arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ss2d_gpu.py -x -q 2>&1 | tail -6 > gpurun_out/r02_ss2d_tests.log; tail -3 gpurun_out/r02_ss2d_tests.log
timeout 1500 python -m pytest tests/test_model_gpu.py -x -q -k "training_step or swapped_modules" 2>&1 | grep -v Warning | tail -8 > gpurun_out/r02_model_train_tests.log; tail -4 gpurun_out/r02_model_train_tests.log
timeout 1500 python tools/model_bench.py train --steps 3 --warmup 2 > gpurun_out/r02_train_swapped.json 2> gpurun_out/r02_train_swapped.err; tail -2 gpurun_out/r02_train_swapped.err
python - <<PY
import json
m=json.load(open("gpurun_out/r02_train_swapped.json"))
for k,v in m["arms"].items(): print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items()})
PY
