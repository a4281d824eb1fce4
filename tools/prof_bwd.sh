set -e
python tools/ncu_one.py configs1 f32 2 > gpurun_out/s10_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_bwd -s 1 -c 1 -o gpurun_out/s10_bwdrp -f python tools/ncu_one.py configs1 f32 2 > gpurun_out/s10_ncu.log 2>&1
tail -2 gpurun_out/s10_ncu.log
