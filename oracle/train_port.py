"""CPU restatement of the Spike2Former TRAINING step's loss path (TEST INFRASTRUCTURE ONLY).

Forward = oracle/port.py with `Ctx(train=True)` (batch-statistics BatchNorm with momentum 0.1, the `quant` surrogate
gradient of surrogate.py:531-538).  This file restates what sits behind it, each function citing the reference lines:

  mmseg MaskFormerHead._seg_data_to_instance_data   mmseg/models/decode_heads/maskformer_head.py:53-106
  ClassificationCost / FocalLossCost / DiceCost     mmdet/models/task_modules/assigners/match_cost.py:199-223, 271-293, 346-395
  HungarianAssigner.assign                          mmdet/models/task_modules/assigners/hungarian_assigner.py:88-145
  MaskPseudoSampler / MaskSamplingResult            mmdet/models/task_modules/samplers/mask_pseudo_sampler.py:28-60
  MaskFormerHead._get_targets_single / _loss_by_feat_single / loss_by_feat
                                                    mmdet/models/dense_heads/maskformer_head.py:295-365, 410-496, 367-408
  CrossEntropyLoss / FocalLoss / DiceLoss           mmdet/models/losses/{cross_entropy_loss.py:12-61,262-301,
                                                    focal_loss.py:12-60, dice_loss.py:10-62,104-146, utils.py:27-73}

Pinned against the reference's OWN loss / matching code executed by oracle/ref_loader.py::load_training
(tests/golden/make_golden_train.py -> golden_train.pt; tests/test_train_oracle_cpu.py re-runs it live when
/root/reference exists).  Config values are the ones of SDTv2_maskformer_DCNpixelDecoder_ade20k.py:94-131.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import port

EPS32 = float(torch.finfo(torch.float32).eps)
# cfg:94-131
LOSS_CFG = dict(cls_weight=1.0, bg_class_weight=0.1, mask_weight=20.0, mask_gamma=2.0, mask_alpha=0.25, dice_weight=1.0,
                dice_eps=1.0, cost_cls=1.0, cost_mask=20.0, cost_dice=1.0, cost_dice_eps=1.0, cost_focal_eps=1e-12)
TRAIN_CFG = dict(assigner=dict(type="mmdet.HungarianAssigner",
                               match_costs=[dict(type="mmdet.ClassificationCost", weight=1.0),
                                            dict(type="mmdet.FocalLossCost", weight=20.0, binary_input=True),
                                            dict(type="mmdet.DiceCost", weight=1.0, pred_act=True, eps=1.0)]),
                 sampler=dict(type="mmdet.MaskPseudoSampler"))


def seg_to_instances(gt_sem_seg, ignore_index=255):
    """decode_heads/maskformer_head.py:77-105: one binary mask per class present in the label map [1, H, W]."""
    classes = torch.unique(gt_sem_seg, sorted=False, return_inverse=False, return_counts=False)
    labels = classes[classes != ignore_index]
    if len(labels) == 0:
        return labels, torch.zeros((0,) + tuple(gt_sem_seg.shape[-2:])).to(gt_sem_seg)
    return labels, torch.stack([gt_sem_seg == c for c in labels]).squeeze(1).long()


def match_cost(cls_score, mask_pred, gt_labels, gt_masks_ds, c=LOSS_CFG):
    """Sum of the three weighted costs, [num_queries, num_gt] (hungarian_assigner.py:122-130)."""
    cost_cls = -cls_score.softmax(-1)[:, gt_labels] * c["cost_cls"]                     # match_cost.py:217-223
    p = mask_pred.flatten(1).sigmoid()                                                   # :271-293
    g = gt_masks_ds.flatten(1).float()
    n = p.shape[1]
    neg = -(1 - p + c["cost_focal_eps"]).log() * (1 - c["mask_alpha"]) * p.pow(c["mask_gamma"])
    pos = -(p + c["cost_focal_eps"]).log() * c["mask_alpha"] * (1 - p).pow(c["mask_gamma"])
    cost_mask = (torch.einsum("nc,mc->nm", pos, g) + torch.einsum("nc,mc->nm", neg, 1 - g)) / n * c["cost_mask"]
    num = 2 * torch.einsum("nc,mc->nm", p, g)                                            # :360-370, pred_act, naive dice
    den = p.sum(-1)[:, None] + g.sum(-1)[None, :]
    cost_dice = (1 - (num + c["cost_dice_eps"]) / (den + c["cost_dice_eps"])) * c["cost_dice"]
    return torch.stack([cost_cls, cost_mask, cost_dice]).sum(dim=0)


def assign(cost):
    """hungarian_assigner.py:132-145 + mask_pseudo_sampler.py:46-60 -> (pos_inds sorted, pos_assigned_gt_inds)."""
    from scipy.optimize import linear_sum_assignment

    rows, cols = linear_sum_assignment(cost.detach().cpu())
    gt_inds = torch.zeros(cost.shape[0], dtype=torch.long)
    gt_inds[torch.from_numpy(rows)] = torch.from_numpy(cols) + 1
    pos = torch.nonzero(gt_inds > 0, as_tuple=False).squeeze(-1).unique()
    return pos, gt_inds[pos] - 1


def weight_reduce_mean(loss, avg_factor):
    """losses/utils.py:57-66 with reduction='mean' and an avg_factor."""
    return loss.sum() / (avg_factor + EPS32)


def loss_single(cls_scores, mask_preds, gt_labels_list, gt_masks_list, num_classes, c=LOSS_CFG):
    """MaskFormerHead._loss_by_feat_single (maskformer_head.py:410-496) for one decoder output.
    cls_scores [B, nq, K+1]; mask_preds [B, nq, h, w]."""
    B, nq = cls_scores.shape[:2]
    labels = torch.full((B, nq), num_classes, dtype=torch.long)
    mask_weights = mask_preds.new_zeros((B, nq))
    targets, avg_factor = [], 0
    for i in range(B):
        gl, gm = gt_labels_list[i], gt_masks_list[i]
        if gm.shape[0] > 0:                                                              # :328-334
            gm_ds = F.interpolate(gm.unsqueeze(1).float(), mask_preds.shape[-2:], mode="nearest").squeeze(1).long()
            pos, pos_gt = assign(match_cost(cls_scores[i], mask_preds[i], gl, gm_ds, c))
        else:
            pos, pos_gt = torch.zeros(0, dtype=torch.long), torch.zeros(0, dtype=torch.long)
        labels[i, pos] = gl[pos_gt]                                                      # :349-353
        mask_weights[i, pos] = 1.0
        targets.append(gm[pos_gt])
        avg_factor += max(pos.numel(), 1)                                                # mask_sampling_result.py:25-28
    mask_targets = torch.cat(targets, dim=0)
    class_weight = cls_scores.new_tensor([1.0] * num_classes + [c["bg_class_weight"]])
    flat_scores, flat_labels = cls_scores.flatten(0, 1), labels.flatten(0, 1)
    ce = F.cross_entropy(flat_scores, flat_labels, weight=class_weight, reduction="none")    # cross_entropy_loss.py:42-47
    loss_cls = c["cls_weight"] * weight_reduce_mean(ce * 1.0, class_weight[flat_labels].sum())  # :450-457
    num_total_masks = max(cls_scores.new_tensor([avg_factor]), 1)                            # :459-460 (reduce_mean: 1 rank)
    mp = mask_preds[mask_weights > 0]
    if mask_targets.shape[0] == 0:                                                           # :467-471
        return loss_cls, mp.sum(), mp.sum()
    mp = F.interpolate(mp.unsqueeze(1), mask_targets.shape[-2:], mode="bilinear", align_corners=False).squeeze(1)
    # dice (dice_loss.py:46-60, naive_dice, eps 1.0, activate -> sigmoid)
    inp, tgt = mp.sigmoid().flatten(1), mask_targets.flatten(1).float()
    a, b, cc = torch.sum(inp * tgt, 1), torch.sum(inp, 1), torch.sum(tgt, 1)
    dice = 1 - (2 * a + c["dice_eps"]) / (b + cc + c["dice_eps"])
    loss_dice = c["dice_weight"] * weight_reduce_mean(dice, num_total_masks)
    # focal on the flattened masks; "target is (1 - mask_targets)": class index 0 = foreground (:486-494)
    h, w = mp.shape[-2:]
    pred = mp.reshape(-1, 1)
    target = F.one_hot(1 - mask_targets.reshape(-1), num_classes=2)[:, :1].type_as(pred)     # focal_loss.py:226-229
    ps = pred.sigmoid()
    pt = (1 - ps) * target + ps * (1 - target)
    fw = (c["mask_alpha"] * target + (1 - c["mask_alpha"]) * (1 - target)) * pt.pow(c["mask_gamma"])
    focal = F.binary_cross_entropy_with_logits(pred, target, reduction="none") * fw
    loss_mask = c["mask_weight"] * weight_reduce_mean(focal, num_total_masks * h * w)
    return loss_cls, loss_mask, loss_dice


def loss_by_feat(all_cls, all_masks, gt_sem_seg, num_classes, ignore_index=255):
    """maskformer_head.py:367-408 on top of the mmseg glue: dict of 7 x 3 loss terms."""
    inst = [seg_to_instances(g, ignore_index) for g in gt_sem_seg]
    gl, gm = [i[0] for i in inst], [i[1] for i in inst]
    out = {}
    L = all_cls.shape[0]
    for l in range(L):
        lc, lm, ld = loss_single(all_cls[l], all_masks[l], gl, gm, num_classes)
        pre = "" if l == L - 1 else f"d{l}."
        out[pre + "loss_cls"], out[pre + "loss_mask"], out[pre + "loss_dice"] = lc, lm, ld
    return out


def train_losses(P, cfg, img, gt_sem_seg):
    """EncoderDecoder.loss (encoder_decoder.py:163-188) in training mode -> dict of loss terms (autograd graph on P)."""
    cx = port.Ctx(P, train=True)
    T = cfg["backbone"]["T"]
    feats = port.backbone_forward(cx, cfg["backbone"], img)
    all_cls, all_masks = port.head_forward(cx, cfg["decode_head"], feats, T)
    return loss_by_feat(all_cls, all_masks, gt_sem_seg, cfg["decode_head"]["num_classes"])


def total_loss(losses):
    """mmengine BaseModel.parse_losses: the sum of every term whose key contains 'loss'."""
    return sum(v for k, v in losses.items() if "loss" in k)
