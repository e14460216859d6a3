"""Host-side parameter preparation: BatchNorm folding, weight re-layout, RepConv re-parameterisation.

All folding is done once per model in float64 and rounded to fp32, then moved to the device.
Reference formulas: eval-mode BatchNorm y = (x - mean) / sqrt(var + eps) * weight + bias (eps 1e-5),
BNAndPadLayer border value (sdtv2.py:64-89), RepConv body (sdtv2.py:111-132).
"""
from __future__ import annotations

import torch

BN_EPS = 1e-5


def bn_affine(sd, key, eps=BN_EPS):
    """-> (s, t) float64 with BN(x) = s*x + t."""
    w, b = sd[key + ".weight"].double(), sd[key + ".bias"].double()
    rm, rv = sd[key + ".running_mean"].double(), sd[key + ".running_var"].double()
    s = w / torch.sqrt(rv + eps)
    return s, b - rm * s


def conv_bn(sd, conv_key, bn_key=None, extra_scale=None):
    """Fold conv bias + BN (+ a per-channel multiplier such as a layer scale) into (scale, shift) fp64."""
    bias = sd.get(conv_key + ".bias")
    cout = sd[conv_key + ".weight"].shape[0]
    if bn_key is not None:
        s, t = bn_affine(sd, bn_key)
    else:
        s, t = torch.ones(cout, dtype=torch.float64), torch.zeros(cout, dtype=torch.float64)
    if bias is not None:
        t = t + bias.double() * s
    if extra_scale is not None:
        e = extra_scale.double()
        s, t = s * e, t * e
    return s, t


def w_khwc(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, kh, kw] (or [Cout, Cin, 1] / [Cout, Cin]) -> [Cout, kh*kw*Cin] with Cin fastest."""
    if w.dim() == 2:
        return w.contiguous()
    if w.dim() == 3:
        return w[:, :, 0].contiguous()
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def dw_taps(w: torch.Tensor) -> torch.Tensor:
    """depthwise [C, 1, k, k] -> tap-major [k*k, C]."""
    c, _, k, _ = w.shape
    return w.reshape(c, k * k).t().contiguous()


def repconv_dense3x3(sd, key):
    """RepConv + outer BN as ONE dense 3x3 convolution with zero padding.

    body = conv1x1(W1) -> BN1 (border padded with BN1(0)) -> dw3x3(D, no pad) -> conv1x1(W2) -> BN2 ; then BN3.
    Because the border value equals BN1 applied to the zero-padded conv1x1 output, the chain is affine in
    the zero-padded input:  out[o,p] = sum_{tap,i} Wm[o,tap,i] x[i,p+tap] + bias[o]  with
      Wm[o,tap,i] = s3 s2 sum_c W2[o,c] D[c,tap] s1[c] W1[c,i]
      bias[o]     = s3 (s2 sum_c W2[o,c] (sum_tap D[c,tap]) t1[c] + t2) + t3
    Returns (Wm [Cout, 9*Cin] fp64 with Cin fastest, bias [Cout] fp64)."""
    W1 = sd[key + ".0.body.0.weight"].double()[:, :, 0, 0]           # [C, Cin]
    s1, t1 = bn_affine(sd, key + ".0.body.1.bn")
    D = sd[key + ".0.body.2.0.weight"].double()[:, 0].reshape(-1, 9)  # [C, 9]
    W2 = sd[key + ".0.body.2.1.weight"].double()[:, :, 0, 0]          # [Cout, C]
    s2, t2 = bn_affine(sd, key + ".0.body.2.2")
    s3, t3 = bn_affine(sd, key + ".1")
    A1 = W1 * s1[:, None]                                            # [C, Cin]
    # Wm[o, tap, i] = sum_c W2[o,c] * D[c,tap] * A1[c,i]
    Wm = torch.einsum("oc,ct,ci->oti", W2, D, A1)
    Wm = Wm * (s3 * s2)[:, None, None]
    bias = s3 * (s2 * (W2 @ (D.sum(1) * t1)) + t2) + t3
    return Wm.reshape(Wm.shape[0], -1), bias


def f32(t, device):
    return t.to(torch.float32).contiguous().to(device)
