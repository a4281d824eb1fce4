#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ss2d_gpu.py tests/test_model_gpu.py -x -q -k "block_tail or swapped_modules or bf16_batch or full_model or graphed" 2>&1 | grep -v Warning | tail -6
python tools/model_bench.py infer > gpurun_out/r02_infer_blocks.json 2> gpurun_out/r02_infer_blocks.err; tail -2 gpurun_out/r02_infer_blocks.err
python - <<PY
import json
m=json.load(open("gpurun_out/r02_infer_blocks.json"))
for k,v in m["arms"].items(): print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items()})
print(m.get("best_arm"), m.get("pairs_per_s"), m.get("e2e"))
PY
python tools/model_bench.py breakdown --arm fused_blocks > gpurun_out/r02_breakdown_fused.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r02_breakdown_fused.json"))
print(d["gpu_time_us"], d["launches"])
for a,b in d["groups"].items(): print(a,b)
for x in d["top"][:14]: print(round(x["us"]), x["count"], x["name"][:90])
PY
