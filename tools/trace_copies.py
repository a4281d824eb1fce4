"""Attribute the ATen copy / elementwise kernels of one fused-blocks inference forward to the Python call sites that launch them
(torch.profiler with stacks): which module code still pays permute / cast / add passes around the library's kernels."""
import collections
import copy
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import model_harness as mh  # noqa: E402

dev = torch.device("cuda", 0)
model = mh.fix_device_attrs(mh.build_model("full", device="cpu", seed=0).eval().to(dev), dev)
mh.set_backend("ours"); mh.set_fuse(None)
mh.swap_ss2d(model); model = mh.fix_device_attrs(model, dev)
mh.swap_layer_norms(model); mh.swap_vss_blocks(model)
x1, x2 = mh.make_pair(int(sys.argv[1]) if len(sys.argv) > 1 else 32, 256, 256, seed=1, device=dev)
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(2):
        model(x1, x2)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
        model(x1, x2)
        torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_time <= 0 or not ev.name.startswith("aten::"):
        continue
    if ev.name not in ("aten::copy_", "aten::add", "aten::add_", "aten::mul", "aten::gelu", "aten::cat", "aten::_to_copy", "aten::clone",
                       "aten::contiguous", "aten::sub", "aten::sigmoid", "aten::silu", "aten::index", "aten::slice_backward"):
        continue
    frames = [f for f in (ev.stack or []) if "site-packages/torch" not in f and "profiler" not in f]
    site = frames[0] if frames else "?"
    key = (ev.name, site.replace(ROOT, "."), str(ev.input_shapes)[:80])
    agg[key][0] += ev.device_time
    agg[key][1] += 1
tot = sum(v[0] for v in agg.values())
print("total attributed us", round(tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(round(v[0]), v[1], k[0], "|", k[1][-110:], "|", k[2])
