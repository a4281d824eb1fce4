"""One fm_block_gates call on a stage-2 sized activation (batch 32, 16x16 tokens, 384 channels) -- ncu / timing target."""
import os, sys, types
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusionmamba_b200 import blocks
B, P, Cc = 32, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 384
torch.manual_seed(0)
x = torch.randn(B, P, Cc, device="cuda").bfloat16()
se = types.SimpleNamespace(global_reduce=torch.nn.Linear(Cc, Cc // 8).cuda(), channel_select=torch.nn.Linear(Cc // 8, Cc).cuda(),
                           norm=torch.nn.LayerNorm(Cc).cuda())
eca = torch.randn(1, 1, 3, device="cuda")
for _ in range(3):
    blocks.block_gates(x, se, eca)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    blocks.block_gates(x, se, eca)
e1.record(); torch.cuda.synchronize()
print("us per call (2 kernels)", e0.elapsed_time(e1) / 50 * 1e3)
