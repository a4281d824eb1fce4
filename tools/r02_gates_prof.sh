#!/bin/bash
mkdir -p gpurun_out
python tools/gates_one.py 384 || exit 1
python tools/gates_one.py 96
ncu --set full --clock-control none --import-source on -k regex:block_gates_finish -s 2 -c 1 -o gpurun_out/r02_gates -f python tools/gates_one.py 384 > gpurun_out/r02_gates_ncu.log 2>&1
tail -2 gpurun_out/r02_gates_ncu.log
