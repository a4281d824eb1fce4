#!/bin/bash
# round-2 first GPU pass: model-level parity + pairs/s + breakdown + training step + 1024^2
mkdir -p gpurun_out
rm -f gpurun_out/model_parity.jsonl
python -m pytest tests/test_model_gpu.py -x -q 2>&1 | grep -v Warning | tail -25 > gpurun_out/r02_model_tests.log
cat gpurun_out/r02_model_tests.log | tail -5
python tools/model_bench.py infer > gpurun_out/r02_model_infer.json 2> gpurun_out/r02_model_infer.err; tail -c 1500 gpurun_out/r02_model_infer.json
for arm in dropin swapped reference_cuda; do
  python tools/model_bench.py breakdown --arm $arm > gpurun_out/r02_breakdown_$arm.json 2> gpurun_out/r02_breakdown_$arm.err
done
python tools/model_bench.py long > gpurun_out/r02_model_long.json 2> gpurun_out/r02_model_long.err; tail -c 800 gpurun_out/r02_model_long.json
python tools/model_bench.py train --steps 2 --warmup 1 > gpurun_out/r02_model_train.json 2> gpurun_out/r02_model_train.err; tail -c 1200 gpurun_out/r02_model_train.json
nvidia-smi --query-gpu=memory.used --format=csv
