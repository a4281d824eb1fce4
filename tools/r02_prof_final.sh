#!/bin/bash
# Final round-2 profiles: (1) launch list of the bench command, (2) ncu --set full of the scan kernels at configs[1] (fp32, bf16) and of
# the pipelined lane-serial backward at the stage-1 shape, each only after the plain command exited 0.  The reports are summarised ON THE BOX
# (tools/ncu_summary.py, tools/ncu_traffic.py) and only the text summaries travel back (the .ncu-rep files exceed the 64 MiB return limit).
mkdir -p gpurun_out /tmp/ncu
python bench.py --steps 2 --warmup 3 --no-model --no-e2e --no-cpu-baseline > gpurun_out/r02_final_plain.json 2> gpurun_out/r02_final_plain.err || { tail gpurun_out/r02_final_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-model --no-e2e --no-cpu-baseline > gpurun_out/r02_final_launches.log 2>&1
cp profiles/ncu_traffic.json /tmp/ncu/before.json
for dt in f32 bf16; do
  python tools/ncu_one.py configs1 $dt 2 > gpurun_out/r02_final_one_$dt.log 2>&1 || { tail gpurun_out/r02_final_one_$dt.log; exit 1; }
  ncu --set full --clock-control none --import-source on -k regex:scan_ -s 2 -c 2 -o /tmp/ncu/final_$dt -f python tools/ncu_one.py configs1 $dt 2 > gpurun_out/r02_final_ncu_$dt.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/final_$dt.ncu-rep > gpurun_out/r02_final_configs1_${dt}_ncu_summary.txt 2>&1
  python tools/ncu_traffic.py configs1_$dt /tmp/ncu/final_$dt.ncu-rep > /dev/null 2>&1
done
cp profiles/ncu_traffic.json gpurun_out/r02_ncu_traffic.json
python tools/ncu_one.py stage1 f32 2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_bwd -s 1 -c 1 -o /tmp/ncu/stage1_bwd -f python tools/ncu_one.py stage1 f32 2 > gpurun_out/r02_final_ncu_stage1.log 2>&1
python tools/ncu_summary.py /tmp/ncu/stage1_bwd.ncu-rep $(python -c "print(32*1536*256*16/32)") > gpurun_out/r02_final_stage1_bwd_ls2_ncu_summary.txt 2>&1
ls -la /tmp/ncu gpurun_out | head -30
