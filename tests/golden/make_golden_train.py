"""Generates tests/golden/golden_train.pt by running THE REFERENCE ITSELF in training mode: backbone + head forward
(batch-statistics BatchNorm, `quant` surrogate gradient), the reference's own CrossEntropy / Focal / Dice losses,
HungarianAssigner and MaskPseudoSampler (oracle/ref_loader.py::load_training executes those files), and autograd.

    python tests/golden/make_golden_train.py

Stored: the 21 loss terms, the gradient of every parameter as (norm, first 8 entries), the updated BatchNorm running
statistics of three layers -- for the tiny config, batch 2, 64x64, seeded labels with an ignored stripe.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import ref_loader, train_port, weights  # noqa: E402
from spike2former_b200 import configs  # noqa: E402


def inputs(cfg, h=64, w=64, batch=2, seed=3):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(batch, 3, h, w, generator=g)
    # blocky label map (8x8 blocks) so that masks are regions, plus an ignored stripe
    coarse = torch.randint(0, cfg["decode_head"]["num_classes"], (batch, 1, h // 8, w // 8), generator=g)
    gt = coarse.repeat_interleave(8, 2).repeat_interleave(8, 3)
    gt[:, :, :4, :] = 255
    return img, gt


def reference_run(cfg, P, img, gt):
    bb, hd = ref_loader.build_reference_for_training(cfg, train_port.TRAIN_CFG)
    bb.load_state_dict({k[9:]: v for k, v in P.items() if k.startswith("backbone.")}, strict=True)
    hd.load_state_dict({k[12:]: v for k, v in P.items() if k.startswith("decode_head.")}, strict=True)
    losses = ref_loader.reference_train_loss(bb, hd, img, gt)
    sum(losses.values()).backward()
    grads = {}
    for prefix, mod in (("backbone.", bb), ("decode_head.", hd)):
        for n, p in mod.named_parameters():
            if p.grad is not None:
                grads[prefix + n] = p.grad.detach().clone()
    stats = {prefix + n: b.detach().clone() for prefix, mod in (("backbone.", bb), ("decode_head.", hd))
             for n, b in mod.named_buffers() if "running_" in n}
    return {k: v.detach().reshape(-1)[0].clone() for k, v in losses.items()}, grads, stats


def main():
    torch.set_num_threads(8)
    cfg = configs.tiny()
    P = weights.calibrated_state(cfg, 64, 64, style="stable")      # bounded flip growth: GPU-vs-CPU rounding does not cascade
    img, gt = inputs(cfg)
    losses, grads, stats = reference_run(cfg, P, img, gt)
    keep_stats = ["backbone.downsample1_1.encode_bn", "backbone.block3.0.attn.q_conv.0.body.1.bn", "decode_head.transformer_decoder.layers.5.ffn.bn2"]
    blob = dict(losses=losses, grad_norm={k: v.norm() for k, v in grads.items()},
                grad_head={k: v.reshape(-1)[:8].clone() for k, v in grads.items()},
                stats={k + s: stats[k + s] for k in keep_stats for s in (".running_mean", ".running_var")})
    torch.save(blob, os.path.join(HERE, "golden_train.pt"))
    print("losses", {k: round(float(v), 5) for k, v in losses.items()})
    print(f"{len(grads)} parameter gradients; total grad norm {float(torch.stack(list(blob['grad_norm'].values())).norm()):.4f}")


if __name__ == "__main__":
    main()
