#!/bin/bash
# ncu --set full of the forward at configs[1] fp32 (dense 8-step checkpoints, as the default backward now wants them); report comes back
mkdir -p gpurun_out
python tools/ncu_one.py configs1 f32 2 > gpurun_out/r02_fwd_plain.log 2>&1 || { tail gpurun_out/r02_fwd_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:scan_fwd -s 1 -c 1 -o gpurun_out/r02_fwd -f python tools/ncu_one.py configs1 f32 2 > gpurun_out/r02_fwd_ncu.log 2>&1
tail -2 gpurun_out/r02_fwd_ncu.log; ls -la gpurun_out/r02_fwd.ncu-rep
