#!/bin/bash
# ncu --set full of the forward and (lane-serial) backward kernels at BASELINE configs[1] fp32; one launch each, after the plain run exited 0
mkdir -p gpurun_out
python tools/ncu_one.py configs1 f32 2 > gpurun_out/r02_prof_plain.log 2>&1 || { tail gpurun_out/r02_prof_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:scan_bwd -s 1 -c 1 -o gpurun_out/r02_bwd_ls -f python tools/ncu_one.py configs1 f32 2 > gpurun_out/r02_ncu_bwd.log 2>&1
tail -2 gpurun_out/r02_ncu_bwd.log
ncu --set full --clock-control none --import-source on -k regex:scan_fwd -s 1 -c 1 -o gpurun_out/r02_fwd16 -f python tools/ncu_one.py configs1 f32 2 > gpurun_out/r02_ncu_fwd.log 2>&1
tail -2 gpurun_out/r02_ncu_fwd.log
ls -la gpurun_out/*.ncu-rep
