set -e
python tools/ncu_one.py configs1 f32 2 > gpurun_out/s55_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_ -s 2 -c 2 -o gpurun_out/s55_final -f python tools/ncu_one.py configs1 f32 2 > gpurun_out/s55_ncu.log 2>&1
tail -2 gpurun_out/s55_ncu.log
