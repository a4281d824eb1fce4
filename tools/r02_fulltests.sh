#!/bin/bash
# the driver's GPU tier: the whole -m gpu suite, then smoke()
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | grep -v Warning | tail -8 ) 2>&1 | tee gpurun_out/r02_fulltests.log | tail -12
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
