"""Synthetic checkpoints: seeded random weights + shipped calibration statistics.

There is no network and no trained checkpoint in this environment, and PyTorch-default init is a
degenerate workload for this network (all spikes zero after the second neuron).
A synthetic checkpoint is therefore made of

  * `random_state(cfg, seed)`: every tensor drawn from one seeded CPU generator in sorted key order:
    conv / linear weights U(-b, b) with b = gain / sqrt(fan_in), BN weight U(1, 2), BN bias U(0.5, 2)
    (zero-mean for the decoder's residual branches), DCN offset/mask convs N(0, 0.05), layer scales 1;
  * calibration statistics (BN running mean / var and three data-dependent biases) that were measured
    once by a training-mode pass of the CPU oracle (tools/make_calibration.py) and are shipped as a small
    file under spike2former_b200/data/.  Loading them needs no oracle code.
"""
from __future__ import annotations

import math
import os

import torch

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

# keys written by the calibration pass besides the BN running statistics
CALIBRATED_EXTRA = ("decode_head.cls_embed.bias", "decode_head.pixel_decoder.mask_feature.bias",
                    "decode_head.pixel_decoder.mask_feature.weight_gain")


def skeleton_state(cfg):
    """Key -> zero tensor of the right shape, from the parameter tree (CPU construction only)."""
    from .models import build_segmentor

    m = build_segmentor(cfg)
    return {k: torch.zeros_like(v) for k, v in m.state_dict().items()}


# ---------------------------------------------------------------------------------------------- "stable" style
# The default style above is CHAOTIC as a dynamical system (tests/test_chaos.py): one flipped spike grows ~20x per
# layer, so two correct fp32 implementations disagree end to end.  The stable style keeps every neuron firing at a
# moderate rate but bounds that growth, which makes FREE-RUNNING parity (no teacher forcing) measurable:
#   * conv / linear rows are heavy-tailed: K_STRONG strong taps per output row over a dense background ~100x
#     smaller (real checkpoints after BN folding look like this; it is also the hard case of the fixed-point packer);
#     depthwise kernels are a strong centre tap + one strong random tap over a weak background;
#   * BatchNorm gains U(0.25, 0.5); the LAST BatchNorm of every residual branch has gain U(0.002, 0.004) and zero-mean
#     bias (branches perturb the stream weakly -- the layer-scale regime the reference itself initialises its
#     encoder with, detr_layers.py:301,329-332 -- and the stream neither drifts nor saturates);
#   * q / k / v BatchNorm biases are chosen from the token count so that the un-normalised attention output
#     (a sum over all keys, no softmax) lands at ATTN_TARGET levels instead of saturating at 8.
K_STRONG, WEAK, DW_WEAK = 2, 0.01, 0.05
GAIN, GAIN_FINAL, ATTN_TARGET = (0.25, 0.5), (0.002, 0.004), 0.6
CLS_GAIN, ME_AMP = 8.0, 6.0      # smooth class scores (one flipped query spike must not switch a whole mask's class)
DAMPED, GAIN_DAMPED = (".output_convs.", ".encoder_out_proj.1."), (0.03, 0.06)   # top-down FPN path: a coarse flip fans out x4 per level
BRANCH_FINAL = (".Conv.bn2.", ".bn2.", ".proj_conv.1.", ".fc2_bn.", ".pwconv2.1.", ".out_conv.1.", ".ffn.bn2.")


def _mean_level(beta, gamma):
    """E[round(clamp(x, 0, 8))] for x ~ N(beta, gamma^2)."""
    return sum(0.5 * (1 + math.erf((beta - (k - 0.5)) / gamma / math.sqrt(2))) for k in range(1, 9))


def _beta_for(level, gamma):
    lo, hi = -10.0, 10.0
    for _ in range(60):
        mid = (lo + hi) / 2
        lo, hi = (mid, hi) if _mean_level(mid, gamma) < level else (lo, mid)
    return (lo + hi) / 2


def stable_state(cfg, seed=1234, hw=(512, 512)):
    g = torch.Generator().manual_seed(seed)
    P = skeleton_state(cfg)
    H, W = hw
    heads = cfg["backbone"]["num_heads"]
    nq = cfg["decode_head"]["num_queries"]
    rnd = lambda shape: 2 * torch.rand(shape, generator=g) - 1
    for k in sorted(P):
        t = P[k]
        leaf = k.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            continue
        if leaf == "running_mean":
            t.zero_()
        elif leaf == "running_var":
            t.fill_(1.0)
        elif leaf in ("gamma1", "gamma2", "gamma3") or k == "decode_head.w":
            t.fill_(1.0)
        elif t.dim() == 1 and (k[: -len(leaf)] + "running_mean") in P:     # BN affine
            final = any(p in k for p in BRANCH_FINAL) and ".dcn.input_proj.pwconv2.1." not in k
            qkv = any(f".attn.{nm}_conv.1." in k for nm in "qkv")
            damped = any(p in k for p in DAMPED)
            if leaf == "weight":
                lo, hi = GAIN_FINAL if final else GAIN_DAMPED if damped else GAIN
                t.copy_(lo + (hi - lo) * torch.rand(t.shape, generator=g))
            elif qkv:
                if k.startswith("backbone."):
                    m = (ATTN_TARGET / (math.sqrt(t.numel() // heads) * (H // 16) * (W // 16))) ** (1 / 3)
                else:
                    li = int(k.split("layers.")[1].split(".")[0])
                    nk = (H // 16) * (W // 16) * 4 ** (li % 3) if ".cross_attn." in k else nq
                    m = (ATTN_TARGET / (2 * nk)) ** (1 / 3)
                t.copy_(_beta_for(8 * m, sum(GAIN) / 2) + 0.3 * rnd(t.shape))
            elif final:
                t.copy_(0.4 * rnd(t.shape))
            elif ".dcn.offset.1." in k:
                t.copy_(0.5 * rnd(t.shape))
            else:
                t.copy_(0.3 + 1.2 * torch.rand(t.shape, generator=g))
        elif "query_embed" in k or "query_feat" in k or "level_embed" in k:
            t.copy_(torch.randn(t.shape, generator=g))
        elif leaf == "weight":
            fan_in = t[0].numel()
            rows = t.shape[0]
            dense = rnd(t.shape)
            if t.dim() == 4 and t.shape[1] == 1 and rows > 1 and fan_in > 1:          # depthwise
                kk = t.shape[-1]
                w = (DW_WEAK * dense).reshape(rows, -1)
                w[:, (kk * kk) // 2] += torch.where(torch.rand(rows, generator=g) < 0.5, -1.0, 1.0)
                idx = torch.randint(0, kk * kk, (rows,), generator=g)
                w[torch.arange(rows), idx] += 0.7 * rnd((rows,))
                t.copy_(w.reshape(t.shape))
            elif fan_in >= 16 and not any(s in k for s in ("cls_embed", "mask_embed.fc_out", "mask_feature", "shortcut")):
                # mask_embed.fc1 / fc2 have no BatchNorm behind them: the strong taps themselves set the firing level
                amp = ME_AMP if "mask_embed" in k else 1.0 / math.sqrt(K_STRONG)
                flat = (WEAK / math.sqrt(fan_in)) * dense.reshape(rows, -1)
                idx = torch.stack([torch.randperm(fan_in, generator=g)[:K_STRONG] for _ in range(rows)])
                flat.scatter_add_(1, idx, rnd((rows, K_STRONG)) * amp)
                t.copy_(flat.reshape(t.shape))
            else:
                gain = CLS_GAIN if "cls_embed" in k else 1.0
                t.copy_(dense * gain / math.sqrt(fan_in))
        elif leaf == "bias":
            t.copy_(rnd(t.shape) / math.sqrt(P[k[: -len("bias")] + "weight"][0].numel()))
        else:  # pragma: no cover
            raise KeyError(f"no init rule for {k}")
    return P


def random_state(cfg, seed=1234, bn_gain=(1.0, 2.0), style="default", hw=(512, 512)):
    if style == "stable":
        return stable_state(cfg, seed, hw)
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


def calibration_of(P):
    """The part of a calibrated state that cannot be regenerated from the seed."""
    out = {k: v.clone() for k, v in P.items() if k.endswith(("running_mean", "running_var"))}
    for k in CALIBRATED_EXTRA[:2]:
        out[k] = P[k].clone()
    return out


def apply_calibration(P, calib):
    for k, v in calib.items():
        if k == CALIBRATED_EXTRA[2]:
            P["decode_head.pixel_decoder.mask_feature.weight"].mul_(float(v))
        else:
            P[k].copy_(v)
    return P


def calibration_path(name):
    return os.path.join(DATA_DIR, f"calib_{name}.pt")


def synthetic_checkpoint(name, cfg, seed=1234):
    """Seeded weights + shipped calibration file `data/calib_<name>.pt` -> full state_dict.
    Names ending in `_stable` (e.g. "ade20k_stable") use the stable style at the shape the file was made for."""
    path = calibration_path(name)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run tools/make_calibration.py in the build container")
    blob = torch.load(path, map_location="cpu")
    if blob["seed"] != seed:
        raise ValueError(f"calibration file was made for seed {blob['seed']}, not {seed}")
    style = blob.get("style", "default")
    return apply_calibration(random_state(cfg, seed, style=style, hw=(blob["h"], blob["w"])), blob["calib"])
