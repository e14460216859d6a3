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
    """Seeded weights + shipped calibration file `data/calib_<name>.pt` -> full state_dict."""
    path = calibration_path(name)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run tools/make_calibration.py in the build container")
    blob = torch.load(path, map_location="cpu")
    if blob["seed"] != seed:
        raise ValueError(f"calibration file was made for seed {blob['seed']}, not {seed}")
    return apply_calibration(random_state(cfg, seed), blob["calib"])
