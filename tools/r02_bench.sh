#!/bin/bash
# full default bench line at N=1 (what the driver runs), summarised
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  ( time python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ) 2>&1 | grep real
else
  ( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err ) 2>&1 | grep real
fi
tail -3 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_n$N.json") if l.startswith("{")][-1])
for k in ("value","ms_per_step","roofline","roofline_fwd","e2e","cpu_baseline","ref_cuda","clocks","gpu_launches","rank_binding"): print(k, json.dumps(d.get(k))[:500])
for k in ("model","train","longseq"): print(k, json.dumps(d.get(k))[:1800])
PY
