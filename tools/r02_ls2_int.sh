#!/bin/bash
# round-2: pipelined lane-serial backward + fused conv/unfold backward integrated -- scan / SS2D parity suites, then the three-kernel A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_scan_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/r02_ls2_tests.log
tail -4 gpurun_out/r02_ls2_tests.log
timeout 1500 python -m pytest tests/test_ss2d_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/r02_ss2d_tests.log
tail -6 gpurun_out/r02_ss2d_tests.log
timeout 900 python tools/r02_ls2_ab.py > gpurun_out/r02_ls2_ab.jsonl 2> gpurun_out/r02_ls2_ab.err; tail -3 gpurun_out/r02_ls2_ab.err
python - <<PY
import json
for l in open("gpurun_out/r02_ls2_ab.jsonl"):
    r=json.loads(l)
    print(r["shape"], r["dtype"], {k:(v.get("fwd_us"),v.get("bwd_us"),v.get("error")) for k,v in r.items() if isinstance(v,dict)})
PY
