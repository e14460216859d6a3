"""Calibrated random initialisation (TEST INFRASTRUCTURE ONLY) -- SURVEY.md section 8d.

PyTorch-default init is a degenerate test for this network (every neuron after
the second one outputs zeros, DCN offset/mask convs are zero, layer scales are
1e-6).  The recipe below gives non-degenerate spikes:

  * every tensor is drawn from one seeded CPU generator, in sorted key order
    (independent of module construction order);
  * conv / linear weights ~ U(-b, b), b = gain/sqrt(fan_in); biases U(-b, b);
  * BN weight ~ U(1, 2), BN bias ~ U(0.5, 2);
  * DCN offset/mask convs ~ N(0, 0.05); encoder layer scales gamma = 1;
  * `mask_embed.fc{1,2}` scaled up (no BN behind them), `cls_embed`/`query_*`
    spread so that the argmax of the logits shows many classes;
  * BN running statistics are then set by ONE training-mode forward of the
    oracle port with momentum 1 (`port.Ctx(calibrate=True)`).
"""
from __future__ import annotations

import math

import torch

from . import port


def skeleton_state(cfg):
    """Key -> zero tensor of the right shape, from the product's parameter tree (CPU construction only)."""
    from spike2former_b200 import build_segmentor

    m = build_segmentor(cfg)
    return {k: torch.zeros_like(v) for k, v in m.state_dict().items()}


def random_state(cfg, seed=1234, bn_gain=(1.0, 2.0)):
    g = torch.Generator().manual_seed(seed)
    P = skeleton_state(cfg)
    for k in sorted(P):
        t = P[k]
        leaf = k.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            continue
        if leaf == "running_mean":
            t.zero_()
        elif leaf == "running_var":
            t.fill_(1.0)
        elif leaf in ("gamma1", "gamma2", "gamma3"):
            t.fill_(1.0)
        elif k == "decode_head.w":
            t.fill_(1.0)
        elif t.dim() == 1 and (k[: -len(leaf)] + "running_mean") in P:     # BN affine
            if leaf == "weight":
                t.copy_(bn_gain[0] + (bn_gain[1] - bn_gain[0]) * torch.rand(t.shape, generator=g))
            elif "transformer_decoder" in k and (".out_conv.1." in k or ".ffn.bn2." in k):
                # zero-mean residual branches keep the query states (and so the class logits) query-specific
                t.copy_(2 * torch.rand(t.shape, generator=g) - 1)
            else:
                t.copy_(0.5 + 1.5 * torch.rand(t.shape, generator=g))
        elif ".dcn.offset.0." in k or ".dcn.mask.0." in k:
            t.copy_(0.05 * torch.randn(t.shape, generator=g))
        elif "query_embed" in k or "query_feat" in k or "level_embed" in k:
            t.copy_(torch.randn(t.shape, generator=g))
        elif leaf == "weight":
            fan_in = t[0].numel()
            gain = 1.0
            if "mask_embed.fc1" in k or "mask_embed.fc2" in k:
                gain = 6.0
            if "cls_embed" in k:
                gain = 48.0
            b = gain / math.sqrt(fan_in)
            t.copy_((2 * torch.rand(t.shape, generator=g) - 1) * b)
        elif leaf == "bias":
            wkey = k[: -len("bias")] + "weight"
            fan_in = P[wkey][0].numel()
            b = 1.0 / math.sqrt(fan_in)
            t.copy_((2 * torch.rand(t.shape, generator=g) - 1) * b)
        else:  # pragma: no cover
            raise KeyError(f"no init rule for {k}")
    return P


def calibration_batch(cfg, h, w, batch=2, seed=4321):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, cfg["backbone"]["in_channels"], h, w, generator=g)


@torch.no_grad()
def calibrate(P, cfg, h, w, batch=2, seed=4321):
    """One training-mode pass with momentum 1: running stats := batch stats (in place in P)."""
    xc = calibration_batch(cfg, h, w, batch, seed)
    port.predict(port.Ctx(P, calibrate=True), cfg, xc)
    # Second (eval) pass: centre the class logits on the mean query state so that the class
    # distribution differs between queries (otherwise one class wins every pixel and the
    # argmax-agreement check is vacuous).
    seen = {}
    want = ("decode_head.decoder_out_spike", "decode_head.pixel_decoder.mask_feature_spike")
    cx = port.Ctx(P, tap=lambda n, x, s: seen.__setitem__(n, s) if n in want else None)
    port.predict(cx, cfg, xc)
    # ... and centre every mask-feature channel over space, so masks are spatially selective.
    sp = seen[want[1]] / 8.0                                              # [n, C, h, w] spikes
    wmf = P["decode_head.pixel_decoder.mask_feature.weight"].mul_(2.0).flatten(1)   # sharper masks
    P["decode_head.pixel_decoder.mask_feature.bias"].copy_(-(wmf @ sp.mean((0, 2, 3))) - 0.014)
    od = 4.0 * seen["decode_head.decoder_out_spike"][-1] / 8.0          # last decoder output, [n, nq, C]
    mean_state = od.reshape(-1, od.shape[-1]).mean(0)
    P["decode_head.cls_embed.bias"].copy_(-(P["decode_head.cls_embed.weight"] @ mean_state))
    return P


def calibrated_state(cfg, h, w, seed=1234, bn_gain=(1.0, 2.0)):
    return calibrate(random_state(cfg, seed, bn_gain), cfg, h, w)


def bn_stats(P):
    return {k: v.clone() for k, v in P.items() if k.endswith(("running_mean", "running_var"))}


def test_image(cfg, h, w, batch=1, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, cfg["backbone"]["in_channels"], h, w, generator=g)
