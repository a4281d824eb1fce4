set -e
export FM_SCAN_FWD16=1
python tools/ncu_one.py stage0 f32 2 > gpurun_out/s8_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_fwd -s 1 -c 1 -o gpurun_out/s8_fwd16_stage0 -f python tools/ncu_one.py stage0 f32 2 > gpurun_out/s8_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_fwd -s 1 -c 1 -o gpurun_out/s8_fwd16_configs1 -f python tools/ncu_one.py configs1 f32 2 >> gpurun_out/s8_ncu.log 2>&1
tail -2 gpurun_out/s8_ncu.log
