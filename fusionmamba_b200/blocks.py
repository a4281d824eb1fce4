"""Inference-side replacements for pieces of the reference's VSSBlock_new that sit right around the SS2D path
("rest of VSSBlock_new"), each a drop-in nn.Module with the reference module's state_dict.

``FastLayerNorm``: nn.LayerNorm over the channel dimension of a channels-last activation (models/cross.py:1334, 1352:
``self.norm`` / ``self.norm2`` of VSSBlock_new, ``BiAttn.norm`` :748, ``ln_1`` / ``ln_2`` of VSSBlock_Cross_new :1273-1274).
At 256x256, batch 32 the model calls it 200 times per forward on (B*H*W, 96 ... 768) fp32 rows; torch's kernel runs at a third
of HBM speed on rows that short (64 us per call at stage 0, 12.8 ms of a 40 ms forward, profiles/r02_breakdown_swapped_first.json).
The row kernel behind ``fm_merge_norm`` (fm_norm.cu: lanes share a row, 128-bit loads, weights in registers) does the same
arithmetic -- biased variance, eps inside the square root, fp32 statistics, fp32 output like torch under autocast -- in one pass.

Under autograd (fp32, as train.py runs) the forward is the same kernel and the backward is ``fm_layer_norm_bwd`` -- one pass
over x and dy instead of ATen's input-gradient kernel plus a gamma/beta column reduction (48 ms of LayerNorm in a 384 ms
training step, profiles/r02_train_breakdown_patched.json).

Anything else (CPU, autocast training, odd channel counts, D > 1024 under autograd) runs nn.LayerNorm's own forward, i.e. the
reference's op -- this is not a fallback of the scan path, which has none.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib, ss2d


class _LayerNormRows(torch.autograd.Function):
    """LayerNorm over the last dimension of a contiguous fp32 (rows, D) view: forward = the row kernel behind fm_merge_norm,
    backward = fm_layer_norm_bwd (one pass over x and dy; statistics recomputed, column sums reduced deterministically)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        D = x.shape[-1]
        shim = _NormShim(weight, bias, eps)
        y = ss2d.merge_norm(x.view(1, -1, D), shim, torch.float32, gate=None, channels_last=True).view(x.shape)
        ctx.save_for_backward(x, weight)
        ctx.eps, ctx.has_bias = eps, bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        D = x.shape[-1]
        rows = x.numel() // D
        gy = gy.contiguous().float()
        dx = torch.empty_like(x)
        dw = torch.empty(D, device=x.device, dtype=torch.float32) if weight is not None else None
        db = torch.empty(D, device=x.device, dtype=torch.float32) if ctx.has_bias else None
        L = _lib.lib()
        need = int(L.fm_layer_norm_bwd_workspace_bytes(D, rows))
        ws = torch.empty(need // 4, device=x.device, dtype=torch.float32)
        q = _lib.FmNormBwdParams()
        q.abi_version, q.dim, q.rows, q.eps = _lib.ABI_VERSION, D, rows, float(ctx.eps)
        q.x, q.dy, q.dx = C.c_void_p(x.data_ptr()), C.c_void_p(gy.data_ptr()), C.c_void_p(dx.data_ptr())
        w32 = weight.detach().float().contiguous() if weight is not None else None
        q.weight = C.c_void_p(w32.data_ptr()) if w32 is not None else None
        q.dweight = C.c_void_p(dw.data_ptr()) if dw is not None else None
        q.dbias = C.c_void_p(db.data_ptr()) if db is not None else None
        q.workspace, q.workspace_bytes = C.c_void_p(ws.data_ptr()), need
        with torch.cuda.device(x.device):
            _lib.check(L.fm_layer_norm_bwd(C.byref(q), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "fm_layer_norm_bwd")
        return dx, (dw.to(weight.dtype) if dw is not None else None), db, None


class _NormShim:
    """what ss2d.merge_norm reads from a norm module"""

    def __init__(self, weight, bias, eps):
        self.weight, self.bias, self.eps = weight, bias, eps


class FastLayerNorm(nn.LayerNorm):
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        D = self.normalized_shape[0] if len(self.normalized_shape) == 1 else -1
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or (self.weight is not None and self.weight.requires_grad))
        if (needs_grad and x.is_cuda and x.dtype == torch.float32 and D == x.shape[-1] and D % 4 == 0 and D <= 1024
                and x.dim() >= 2 and x.numel() > 0 and not torch.is_autocast_enabled("cuda")):
            # training (fp32, like train.py): both directions on this library's kernels
            return _LayerNormRows.apply(x.contiguous(), self.weight, self.bias, self.eps)
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
