"""Inference-side replacements for pieces of the reference's VSSBlock_new that sit right around the SS2D path
("rest of VSSBlock_new"), each a drop-in nn.Module with the reference module's state_dict.

``FastLayerNorm``: nn.LayerNorm over the channel dimension of a channels-last activation (models/cross.py:1334, 1352:
``self.norm`` / ``self.norm2`` of VSSBlock_new, ``BiAttn.norm`` :748, ``ln_1`` / ``ln_2`` of VSSBlock_Cross_new :1273-1274).
At 256x256, batch 32 the model calls it 200 times per forward on (B*H*W, 96 ... 768) fp32 rows; torch's kernel runs at a third
of HBM speed on rows that short (64 us per call at stage 0, 12.8 ms of a 40 ms forward, profiles/r02_breakdown_swapped_first.json).
The row kernel behind ``fm_merge_norm`` (fm_norm.cu: lanes share a row, 128-bit loads, weights in registers) does the same
arithmetic -- biased variance, eps inside the square root, fp32 statistics, fp32 output like torch under autocast -- in one pass.

Only the no-grad CUDA fp32 path is replaced; anything else (training, CPU, other dtypes, odd channel counts) runs
nn.LayerNorm's own forward, i.e. the reference's op -- this is not a fallback of the scan path, which has none.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ss2d


class FastLayerNorm(nn.LayerNorm):
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        D = self.normalized_shape[0] if len(self.normalized_shape) == 1 else -1
        if (x.is_cuda and x.dtype in (torch.float32, torch.bfloat16, torch.float16) and D == x.shape[-1] and D % 4 == 0
                and D <= 2048 and x.dim() >= 2 and x.numel() > 0 and x.numel() // D < 2 ** 31
                and not (torch.is_grad_enabled() and (x.requires_grad or (self.weight is not None and self.weight.requires_grad)))):
            # 16-bit activations (outputs of autocast linears) are normalised in fp32 with an fp32 result when autocast is on,
            # exactly what torch's autocast policy for layer_norm does; outside autocast the result keeps the input dtype
            odt = torch.float32 if (x.dtype == torch.float32 or torch.is_autocast_enabled("cuda")) else x.dtype
            xf = x.float().contiguous()              # no-ops for the fp32 channels-last residual stream
            y = ss2d.merge_norm(xf.view(1, -1, D), self, odt, gate=None, channels_last=True)
            return y.view(x.shape)
        return super().forward(x)


def adopt_layer_norms(model: nn.Module) -> int:
    """Opt-in, harness-level: replace every plain ``nn.LayerNorm`` over one (last) dimension inside ``model`` by
    ``FastLayerNorm`` sharing the same parameters (state_dict keys unchanged).  Returns the number replaced."""
    n = 0
    for parent in list(model.modules()):
        for cname, child in list(parent.named_children()):
            if type(child) is nn.LayerNorm and len(child.normalized_shape) == 1:
                new = FastLayerNorm(child.normalized_shape, eps=child.eps, elementwise_affine=child.elementwise_affine,
                                    bias=child.bias is not None)
                new.weight, new.bias = child.weight, child.bias
                new.train(child.training)
                setattr(parent, cname, new)
                n += 1
    return n
