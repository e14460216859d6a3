"""Drop-in for the reference's neuron, `Q_IFNode(surrogate_function=Quant())`
(Segmentation/Qtrick_architecture/clock_driven/neuron.py:395-550, surrogate.py:522-538, base.py:54-69).

Forward:  v = v + x ; s = round_half_even(clamp(v, 0, 8)) ; v = v - s ; return s / 8      (one fused kernel, s2f_nilif_fwd)
Backward: d out / d x = 1/8 on 0 <= v + x <= 8, else 0 -- `quant.backward` followed by the "/ 8"    (s2f_nilif_bwd)

Same constructor keywords, `.v`, `.reset()` and `functional.reset_net` duck typing as the reference, so the reference's
own model files (and ResetModelHook, resetmodel_hook.py:17-37) can use it unchanged on CUDA tensors.  The membrane `v`
is carried between calls until `reset()` exactly like `MemoryModule`; gradients flow through the spike (STE), the soft
reset term is detached (it has no consumer at T = 1).  CPU tensors raise: there is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .registry import register_everywhere


class _NiLifFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, v_in, d_max, norm):
        levels, v_out, y = ops.nilif(x, v_in=v_in, want_v_out=True, want_norm=True, d_max=d_max, norm=norm)
        pre = x if v_in is None else x + v_in                       # what `quant` saw: the charged membrane
        ctx.save_for_backward(pre)
        ctx.d_max, ctx.norm = d_max, norm
        ctx.mark_non_differentiable(v_out)
        return y, v_out

    @staticmethod
    def backward(ctx, gy, _gv):
        (pre,) = ctx.saved_tensors
        gx = ops.nilif_bwd(pre, gy.contiguous(), d_max=ctx.d_max, norm=ctx.norm)
        return gx, None, None, None


@register_everywhere
class Q_IFNode(nn.Module):
    def __init__(self, v_threshold: float = 1.0, v_reset: float = 0.0, surrogate_function=None, detach_reset: bool = False,
                 cupy_fp32_inference: bool = False, d_max: float = ops.D_MAX, norm: float = ops.NORM):
        super().__init__()
        self.v_threshold, self.v_reset, self.detach_reset = v_threshold, v_reset, detach_reset
        self.surrogate_function = surrogate_function                # kept for config compatibility; the kernel is Quant()
        self.d_max, self.norm = float(d_max), float(norm)
        self.v = 0.0                                                # python float after reset(), like MemoryModule

    def reset(self):
        self.v = 0.0

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("Q_IFNode: CUDA tensor required -- spike2former_b200 has no CPU path")
        x = x.contiguous().float()
        v_in = None if isinstance(self.v, float) else self.v
        if isinstance(self.v, float) and self.v != 0.0:
            v_in = torch.full_like(x, self.v)
        y, v_out = _NiLifFn.apply(x, v_in, self.d_max, self.norm)
        self.v = v_out.detach()
        return y
