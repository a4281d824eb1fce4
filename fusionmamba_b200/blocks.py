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
import torch.nn.functional as F
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


# ---------------------------------------------------------------------------------------------------------------------------
# The inference tail of VSSBlock_new (ECA, LDC conv, BiAttn x2, residual adds, norm2): kernels of fm_block.cu
# ---------------------------------------------------------------------------------------------------------------------------
_DT = {torch.float32: _lib.FM_F32, torch.float16: _lib.FM_F16, torch.bfloat16: _lib.FM_BF16}


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def block_gates(x: torch.Tensor, se: nn.Module, eca_weight=None):
    """x (B, P, C) channels-last activation -> (eca_scale or None, se_gate), both (B, C) fp32: the per-(batch, channel) gates of
    eca_layer (models/cross.py:1236-1259) and BiAttn (:744-768) from ONE pass over x (C ABI: fm_block_gates)."""
    B, P, Cc = x.shape
    L = _lib.lib()
    q = _lib.FmBlockGatesParams()
    q.abi_version, q.dtype = _lib.ABI_VERSION, _DT[x.dtype]
    q.batch, q.positions, q.dim, q.reduce_dim, q.eps = B, P, Cc, se.global_reduce.out_features, float(se.norm.eps)
    need = int(L.fm_block_gates_workspace_bytes(B, P, Cc))
    ws = torch.empty(need // 4, device=x.device, dtype=torch.float32)
    gate = torch.empty(B, Cc, device=x.device, dtype=torch.float32)
    scale = torch.empty(B, Cc, device=x.device, dtype=torch.float32) if eca_weight is not None else None
    p_ = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    q.x, q.ln_weight, q.ln_bias = p_(x), p_(se.norm.weight), p_(se.norm.bias)
    q.eca_weight = p_(eca_weight)
    q.w1, q.b1, q.w2, q.b2 = p_(se.global_reduce.weight), p_(se.global_reduce.bias), p_(se.channel_select.weight), p_(se.channel_select.bias)
    q.eca_scale, q.se_gate, q.workspace, q.workspace_bytes = p_(scale), p_(gate), p_(ws), need
    with torch.cuda.device(x.device):
        _lib.check(L.fm_block_gates(C.byref(q), _stream(x.device)), "fm_block_gates")
    return scale, gate


def block_scale(x: torch.Tensor, gate: torch.Tensor) -> torch.Tensor:
    """y = x + x * gate[b, c] (C ABI: fm_block_scale): the ECA apply and the add that feeds the LDC conv (models/cross.py:1365-1369)."""
    B, P, Cc = x.shape[0], x.numel() // (x.shape[0] * x.shape[-1]), x.shape[-1]
    y = torch.empty_like(x)
    q = _lib.FmBlockScaleParams()
    q.abi_version, q.dtype, q.batch, q.positions, q.dim = _lib.ABI_VERSION, _DT[x.dtype], B, P, Cc
    q.x, q.gate, q.y = C.c_void_p(x.data_ptr()), C.c_void_p(gate.data_ptr()), C.c_void_p(y.data_ptr())
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().fm_block_scale(C.byref(q), _stream(x.device)), "fm_block_scale")
    return y


def block_combine_norm(inp: torch.Tensor, x_ssm: torch.Tensor, x_conv: torch.Tensor, g1: torch.Tensor, g2: torch.Tensor, norm2):
    """x' = inp + (x_ssm * g1 + x_conv * g2) in fp32 and LayerNorm(x') in the activation dtype (C ABI: fm_block_combine_norm):
    both BiAttn applies, their sum, the residual add and norm2 of models/cross.py:1370-1375 in one row pass."""
    B, Cc = inp.shape[0], inp.shape[-1]
    P = inp.numel() // (B * Cc)
    x_out = torch.empty_like(inp)
    y_out = torch.empty_like(x_ssm)
    q = _lib.FmBlockCombineParams()
    q.abi_version, q.dtype, q.batch, q.positions, q.dim = _lib.ABI_VERSION, _DT[x_ssm.dtype], B, P, Cc
    q.input_dtype = _DT[inp.dtype]
    q.eps = float(norm2.eps) if norm2 is not None else 1e-5
    p_ = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    q.input, q.x_ssm, q.x_conv, q.gate_ssm, q.gate_conv = p_(inp), p_(x_ssm), p_(x_conv), p_(g1), p_(g2)
    q.ln_weight = p_(norm2.weight) if norm2 is not None else None
    q.ln_bias = p_(norm2.bias) if norm2 is not None else None
    q.x_out, q.y_out = p_(x_out), p_(y_out)
    with torch.cuda.device(inp.device):
        _lib.check(_lib.lib().fm_block_combine_norm(C.byref(q), _stream(inp.device)), "fm_block_combine_norm")
    return x_out, y_out


def norm_lowp(norm: nn.LayerNorm, x: torch.Tensor):
    """LayerNorm(x) written straight in the autocast dtype by the norm-only form of fm_block_combine_norm, or None when that does
    not apply.  Under autocast torch's layer_norm returns fp32 and the Linear that consumes it casts to the autocast dtype: the
    same values (one rounding of the fp32 result) through three passes (upcast copy of a 16-bit stream, fp32 result, downcast copy)
    instead of one.  Only for call sites whose sole consumer is an autocast Linear (the block's first norm -> SS2D.in_proj)."""
    if not (x.is_cuda and torch.is_autocast_enabled("cuda") and not torch.is_grad_enabled() and isinstance(norm, nn.LayerNorm)
            and len(norm.normalized_shape) == 1 and norm.normalized_shape[0] == x.shape[-1] and x.is_contiguous() and x.dim() >= 2):
        return None
    dt = torch.get_autocast_dtype("cuda")
    Cc = x.shape[-1]
    if dt not in _DT or x.dtype not in (torch.float32, dt) or Cc % 4 or Cc > 1024 or x.numel() == 0 or x.shape[0] > 65535:
        return None
    B = x.shape[0]
    y = torch.empty(x.shape, device=x.device, dtype=dt)
    q = _lib.FmBlockCombineParams()
    q.abi_version, q.dtype, q.batch, q.positions, q.dim = _lib.ABI_VERSION, _DT[dt], B, x.numel() // (B * Cc), Cc
    q.input_dtype, q.eps = _DT[x.dtype], float(norm.eps)
    p_ = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    q.input, q.ln_weight, q.ln_bias, q.y_out = p_(x), p_(norm.weight), p_(norm.bias), p_(y)
    q.x_ssm = q.x_conv = q.gate_ssm = q.gate_conv = q.x_out = None
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().fm_block_combine_norm(C.byref(q), _stream(x.device)), "fm_block_combine_norm")
    return y


def _ldc_weight(ldc: nn.Module, dtype: torch.dtype) -> torch.Tensor:
    """The LDC conv's effective weight conv.weight * mask (models/cross.py:807-810), which the reference rebuilds with six small
    kernels on every forward, cached in the activation dtype on the parameters' version counters."""
    srcs = (ldc.conv.weight, ldc.base_mask, ldc.learnable_mask, ldc.learnable_theta)
    key = tuple((t.data_ptr(), t._version) for t in srcs) + (dtype,)
    ent = ldc.__dict__.get("_fm_weight")
    if ent is None or ent[0] != key:
        with torch.no_grad(), torch.autocast("cuda", enabled=False):
            w = ldc.conv.weight
            mask = ldc.base_mask - ldc.learnable_theta * ldc.learnable_mask[:, :, None, None] * \
                ldc.center_mask.to(w.device) * w.sum(2).sum(2)[:, :, None, None]
            ent = ldc.__dict__["_fm_weight"] = (key, (w * mask).to(dtype).contiguous(memory_format=torch.channels_last))
    return ent[1]


def _lowp_param(mod: nn.Module, name: str, dtype: torch.dtype):
    """``getattr(mod, name)`` in ``dtype``, cached on the parameter's storage and version counter (the same rule as SS2D's inference
    cache, ss2d.SS2D._cached): under autocast torch re-casts every fp32 weight and bias on every forward -- four few-microsecond
    kernels per Mlp, 212 launches per forward of the full model (profiles/r02_breakdown_fused_blocks.json)."""
    t = getattr(mod, name)
    if t is None or t.dtype == dtype:
        return t
    key = (t.data_ptr(), t._version, dtype)
    cache = mod.__dict__.setdefault("_fm_lowp", {})
    ent = cache.get(name)
    if ent is None or ent[0] != key:
        with torch.no_grad():
            ent = cache[name] = (key, t.detach().to(dtype))
    return ent[1]


def _mlp_forward(mlp: nn.Module, y: torch.Tensor) -> torch.Tensor:
    """Inference forward of the reference's Mlp (models/cross.py:770-788: fc1 -> GELU -> fc2, dropout inactive) with the autocast
    copies of its weights cached; anything else runs the module itself."""
    fc1, fc2 = getattr(mlp, "fc1", None), getattr(mlp, "fc2", None)
    ok = (isinstance(fc1, nn.Linear) and isinstance(fc2, nn.Linear) and isinstance(mlp.act, nn.GELU)
          and getattr(mlp.act, "approximate", "none") == "none" and not (mlp.training and mlp.drop.p > 0)
          and y.is_cuda and torch.is_autocast_enabled("cuda") and y.dtype == torch.get_autocast_dtype("cuda"))
    if not ok:
        return mlp(y)
    dt = y.dtype
    h = F.linear(y, _lowp_param(fc1, "weight", dt), _lowp_param(fc1, "bias", dt))
    return F.linear(F.gelu(h), _lowp_param(fc2, "weight", dt), _lowp_param(fc2, "bias", dt))


def _vss_fast_ok(blk, x: torch.Tensor) -> bool:
    if not (x.is_cuda and x.dtype in _DT and x.dim() == 4 and x.is_contiguous()):
        return False
    if torch.is_grad_enabled() and (x.requires_grad or any(p_.requires_grad for p_ in blk.parameters())):
        return False
    Cc = x.shape[-1]
    se, eca, ldc = blk.se, blk.self_attention_cross_channel, blk.conv_branch
    dp = getattr(blk.drop_path, "drop_prob", 0.0)
    return (Cc % 4 == 0 and Cc <= 1024 and x.shape[0] <= 65535 and (not blk.training or not dp)
            and isinstance(se.act_fn, nn.GELU) and getattr(se.act_fn, "approximate", "none") == "none"
            and isinstance(se.gate_fn, nn.Sigmoid) and isinstance(se.norm, nn.LayerNorm) and se.norm.elementwise_affine
            and isinstance(eca.conv, nn.Conv1d) and eca.conv.kernel_size == (3,) and eca.conv.padding == (1,) and eca.conv.bias is None
            and isinstance(ldc.conv, nn.Conv2d) and next(blk.parameters()).dtype == torch.float32
            and (not blk.mlp_branch or isinstance(blk.norm2, nn.LayerNorm)))


def _vss_block_forward(self, input: torch.Tensor) -> torch.Tensor:
    """Inference forward of a reference VSSBlock_new (models/cross.py:1362-1377) with its tail on fm_block.cu; bound onto the
    reference's block instances by ``adopt_vss_blocks``.  Anything the fast path does not cover runs the block's own forward."""
    if not _vss_fast_ok(self, input):
        return self._fm_orig_forward(input)
    B, H, W, Cc = input.shape
    xn = norm_lowp(self.norm, input)                                               # LN1 straight into the autocast dtype
    x_ssm = self.op(xn if xn is not None else self.norm(input)).contiguous()       # + SS2D
    if x_ssm.dtype not in _DT or input.dtype not in (torch.float32, x_ssm.dtype):
        return self._fm_orig_forward(input)
    eca_scale, g1 = block_gates(x_ssm.view(B, H * W, Cc), self.se, self.self_attention_cross_channel.conv.weight)
    xin = block_scale(x_ssm, eca_scale)                                            # x_ssm + ECA(x_ssm)
    ldc = self.conv_branch
    with torch.autocast("cuda", enabled=False):
        bias = ldc.conv.bias.to(xin.dtype) if ldc.conv.bias is not None else None
        xc = F.conv2d(xin.permute(0, 3, 1, 2), _ldc_weight(ldc, xin.dtype), bias, ldc.conv.stride, ldc.conv.padding,
                      ldc.conv.dilation, ldc.conv.groups)
    x_conv = xc.permute(0, 2, 3, 1).contiguous()                                   # no copy when cuDNN answers channels-last
    _, g2 = block_gates(x_conv.view(B, H * W, Cc), self.se, None)
    x_new, y2 = block_combine_norm(input, x_ssm, x_conv, g1, g2, self.norm2 if self.mlp_branch else None)
    if not self.mlp_branch:
        return x_new
    return x_new + _mlp_forward(self.mlp, y2)


def adopt_vss_blocks(model: nn.Module) -> int:
    """Opt-in, harness-level: give every reference ``VSSBlock_new`` inside ``model`` the fused inference forward above (an
    instance-level override: parameters, buffers and state_dict keys are untouched).  Returns the number of blocks adopted."""
    import types
    n = 0
    for m in model.modules():
        if type(m).__name__ == "VSSBlock_new" and all(hasattr(m, a) for a in
                                                      ("norm", "op", "conv_branch", "self_attention_cross_channel", "se", "drop_path")):
            if "_fm_orig_forward" not in m.__dict__:
                m.__dict__["_fm_orig_forward"] = m.forward
                m.forward = types.MethodType(_vss_block_forward, m)
                n += 1
    return n
