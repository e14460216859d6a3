"""Surrogate-gradient training step (BASELINE.json config 5).

    EncoderDecoder.loss            mmseg/models/segmentors/encoder_decoder.py:163-188
    mmseg MaskFormerHead.loss      mmseg/models/decode_heads/maskformer_head.py:53-136
    mmdet MaskFormerHead.loss_by_feat / _loss_by_feat_single / _get_targets_single
                                   mmdet/models/dense_heads/maskformer_head.py:367-496, 295-365
    HungarianAssigner + costs      mmdet/models/task_modules/assigners/{hungarian_assigner.py:88-145, match_cost.py}
    optimizer                      configs/Spike2Former/SDTv2_maskformer_DCNpixelDecoder_ade20k.py:137-154
    gradient all-reduce            tools/dist_train.sh:5-11 (torch DDP over NCCL in the reference)

What runs where:
  * every neuron is the hand-written training kernel pair s2f_nilif_train_fwd / s2f_nilif_train_bwd: the forward emits
    the normalised spikes plus ONE tag byte per neuron (level + clamp flag), the backward is `quant.backward`
    (surrogate.py:531-538) from that byte -- the fp32 pre-activation (4 B x 188 M neurons per image) is never saved;
  * convolutions / linears / BatchNorm (batch statistics, momentum 0.1) and their dgrad / wgrad run on cuDNN / cuBLAS
    through torch autograd (plain library GEMMs; fp32 by default, `precision="tf32" | "bf16"` for throughput), the DCNv3
    sampler is `F.grid_sample` exactly as the reference trains it (dcnv3_func.py:147-189);
  * matching: all cost matrices of the 7 decoder outputs x B images are computed on the device, leave it in ONE
    device->host copy, scipy solves them on the host, the assignments return in one host->device copy (the reference
    synchronises 7 x B times per step, hungarian_assigner.py:125-131);
  * data parallelism: one process per GPU, gradients are all-reduced over NCCL in ~25 MB buckets that are launched from
    autograd hooks while the rest of the backward pass is still running; the 7 `reduce_mean(num_total_masks)` scalars
    (maskformer_head.py:459) travel as one 7-vector.
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import ops

BN_EPS, BN_MOMENTUM = 1e-5, 0.1


# ------------------------------------------------------------------------------------------------ primitives
class _Lif(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y, tag = ops.nilif_train_fwd(x)
        ctx.save_for_backward(tag)
        return y

    @staticmethod
    def backward(ctx, gy):
        (tag,) = ctx.saved_tensors
        return ops.nilif_train_bwd(tag, gy.contiguous())


def lif(x):
    """Q_IFNode(Quant()) after reset (neuron.py:166-197): spikes / 8 with the surrogate gradient."""
    if not x.is_cuda:
        raise RuntimeError("spike2former_b200.train: CUDA tensor required (no CPU path)")
    return _Lif.apply(x.contiguous().float())


class Net:
    """Functional training forward over the live parameters of an EncoderDecoder (keys as in its state_dict)."""

    def __init__(self, seg):
        self.P = {**dict(seg.named_parameters()), **dict(seg.named_buffers())}
        self.bb, self.hd = seg.backbone, seg.decode_head
        self.pe = {}

    # -- layers
    def bn(self, key, x):
        P = self.P
        return F.batch_norm(x, P[key + ".running_mean"], P[key + ".running_var"], P[key + ".weight"], P[key + ".bias"],
                            training=True, momentum=BN_MOMENTUM, eps=BN_EPS)

    def conv(self, key, x, stride=1, pad=0, groups=1):
        w, b = self.P[key + ".weight"], self.P.get(key + ".bias")
        if w.dim() == 3:
            return F.conv1d(x, w, b)
        return F.conv2d(x, w, b, stride=stride, padding=pad, groups=groups)

    # -- backbone (sdtv2.py)
    def rep_conv(self, key, x):
        """RepConv + outer BN (sdtv2.py:111-132, 48-89, 280-296); border = BN(0) from the running statistics, detached."""
        P = self.P
        y = self.conv(key + ".0.body.0", x)
        k = key + ".0.body.1.bn"
        y = self.bn(k, y)
        pv = (P[k + ".bias"].detach() - P[k + ".running_mean"] * P[k + ".weight"].detach() /
              torch.sqrt(P[k + ".running_var"] + BN_EPS)).view(1, -1, 1, 1)
        y = F.pad(y, [1, 1, 1, 1])
        y[:, :, 0:1, :] = pv
        y[:, :, -1:, :] = pv
        y[:, :, :, 0:1] = pv
        y[:, :, :, -1:] = pv
        y = self.conv(key + ".0.body.2.0", y, groups=y.shape[1])
        y = self.conv(key + ".0.body.2.1", y)
        return self.bn(key + ".1", self.bn(key + ".0.body.2.2", y))

    def downsample(self, key, x, stride, pad, first):
        if not first:
            x = lif(x)
        return self.bn(key + ".encode_bn", self.conv(key + ".encode_conv", x, stride, pad))

    def sep_conv(self, key, x):
        x = self.bn(key + ".bn1", self.conv(key + ".pwconv1", lif(x)))
        x = self.conv(key + ".dwconv", lif(x), pad=3, groups=x.shape[1])
        return self.bn(key + ".bn2", self.conv(key + ".pwconv2", x))

    def conv_block(self, key, x):
        x = self.sep_conv(key + ".Conv", x) + x
        y = self.bn(key + ".bn1", self.conv(key + ".conv1", lif(x), pad=1))
        return x + self.bn(key + ".bn2", self.conv(key + ".conv2", lif(y), pad=1))

    def sdsa(self, key, x, heads):
        n, c, h, w = x.shape
        tok, d = h * w, c // heads
        x = lif(x)

        def split(t):
            t = lif(t).flatten(2)
            return t.transpose(-1, -2).reshape(n, tok, heads, d).permute(0, 2, 1, 3)

        q, k, v = (split(self.rep_conv(key + f".{nm}_conv", x)) for nm in "qkv")
        y = (q @ (k.transpose(-2, -1) @ v)) * d ** -0.5                                 # sdtv2.py:335-336
        y = lif(y.transpose(2, 3).reshape(n, c, tok))
        return self.rep_conv(key + ".proj_conv", y.reshape(n, c, h, w))

    def ms_mlp(self, key, x):
        n, c, h, w = x.shape
        x = self.bn(key + ".fc1_bn", self.conv(key + ".fc1_conv", lif(x.flatten(2))))
        x = self.bn(key + ".fc2_bn", self.conv(key + ".fc2_conv", lif(x)))
        return x.reshape(n, c, h, w)

    def ms_block(self, key, x, heads):
        x = x + self.sdsa(key + ".attn", x, heads)
        return x + self.ms_mlp(key + ".mlp", x)

    def backbone(self, img):
        bb, b = self.bb, "backbone."
        x = img.unsqueeze(0).repeat(bb.T, 1, 1, 1, 1).flatten(0, 1)
        x = self.downsample(b + "downsample1_1", x, 2, 3, True)
        x1 = x = self.conv_block(b + "ConvBlock1_1.0", x)
        x = self.downsample(b + "downsample1_2", x, 2, 1, False)
        x2 = x = self.conv_block(b + "ConvBlock1_2.0", x)
        x = self.downsample(b + "downsample2", x, 2, 1, False)
        x = self.conv_block(b + "ConvBlock2_1.0", x)
        x3 = x = self.conv_block(b + "ConvBlock2_2.0", x)
        x = self.downsample(b + "downsample3", x, 2, 1, False)
        for j in range(6):
            x = self.ms_block(b + f"block3.{j}", x, bb.num_heads)
        x = self.downsample(b + "downsample4", x, 1, 1, False)
        for j in range(2):
            x = self.ms_block(b + f"block4.{j}", x, bb.num_heads)
        return [x1, x2, x3, x]

    # -- pixel decoder (pixel_decoder.py, detr_layers.py, SNN_core.py, dcnv3.py)
    def sepconv_spike(self, key, x, k):
        x = x.permute(0, 3, 1, 2)
        x = self.bn(key + ".pwconv1.1", self.conv(key + ".pwconv1.0", lif(x)))
        x = self.bn(key + ".dwconv.1", self.conv(key + ".dwconv.0", lif(x), pad=(k - 1) // 2, groups=x.shape[1]))
        x = self.bn(key + ".pwconv2.1", self.conv(key + ".pwconv2.0", lif(x)))
        return x.permute(0, 2, 3, 1)

    def dcn_constants(self, hin, win, hout, wout, ksz, group, offset_scale, dev):
        """_get_reference_points + _generate_dilation_grids (dcnv3_func.py:91-144): depend on the shape only, so they
        are built once (on the host, as the reference does) and kept on the device -- nothing inside the forward copies
        from host memory, which keeps the whole forward CUDA-graph capturable."""
        key = (hin, win, hout, wout, ksz, group, offset_scale, str(dev))
        if key not in self.pe:
            half = (ksz - 1) // 2
            ry = torch.linspace(half + 0.5, half + 0.5 + (hout - 1), hout, dtype=torch.float32)
            rx = torch.linspace(half + 0.5, half + 0.5 + (wout - 1), wout, dtype=torch.float32)
            ref_y, ref_x = torch.meshgrid(ry, rx, indexing="ij")
            ref = torch.stack((ref_x.reshape(-1)[None] / win, ref_y.reshape(-1)[None] / hin), -1).reshape(1, hout, wout, 1, 2)
            lin = torch.linspace(-half, -half + (ksz - 1), ksz, dtype=torch.float32)
            gx, gy = torch.meshgrid(lin, lin, indexing="ij")
            grid = torch.stack([gx / win, gy / hin], -1).reshape(-1, 1, 2).repeat(1, group, 1).permute(1, 0, 2)
            grid = grid.reshape(1, 1, 1, group * ksz * ksz, 2)
            norm = torch.tensor([win, hin]).reshape(1, 1, 1, 2).repeat(1, 1, 1, group * ksz * ksz)
            self.pe[key] = ((ref + grid * offset_scale).to(dev), norm.to(dev))
        return self.pe[key]

    def dcnv3_core(self, x, offset, mask, ksz, pad, group, gch, offset_scale):
        """dcnv3_core_pytorch (dcnv3_func.py:147-189), stride 1, dilation 1."""
        x = F.pad(x, [0, 0, pad, pad, pad, pad])
        n, hin, win, _ = x.shape
        _, hout, wout, _ = offset.shape
        base, norm = self.dcn_constants(hin, win, hout, wout, ksz, group, offset_scale, x.device)
        loc = base.repeat(n, 1, 1, 1, 1).flatten(3, 4) + offset * offset_scale / norm
        pts = ksz * ksz
        xin = x.view(n, hin * win, group * gch).transpose(1, 2).reshape(n * group, gch, hin, win)
        sg = (2 * loc - 1).view(n, hout * wout, group, pts, 2).transpose(1, 2).flatten(0, 1)
        samp = F.grid_sample(xin, sg, mode="bilinear", padding_mode="zeros", align_corners=False)
        m = mask.view(n, hout * wout, group, pts).transpose(1, 2).reshape(n * group, 1, hout * wout, pts)
        out = (samp * m).sum(-1).view(n, group * gch, hout * wout)
        return out.transpose(1, 2).reshape(n, hout, wout, -1)

    def dcn(self, key, inp, group, dw_k):
        n, h, w, c = inp.shape
        x = self.sepconv_spike(key + ".input_proj", inp, dw_k)
        x1 = lif(inp.permute(0, 3, 1, 2))
        x1 = lif(self.bn(key + ".dw_conv.1", self.conv(key + ".dw_conv.0", x1, pad=(dw_k - 1) // 2, groups=c)))
        offset = self.bn(key + ".offset.1", self.conv(key + ".offset.0", x1)).reshape(n, h, w, -1)      # reinterpreting reshapes
        mask = lif(self.bn(key + ".mask.1", self.conv(key + ".mask.0", x1)).reshape(n, h, w, -1))       # dcnv3.py:214-215
        x = self.dcnv3_core(x.contiguous(), offset, mask, 3, 1, group, c // group, 1.0)
        return self.sepconv_spike(key + ".output_proj", x, dw_k)

    def enc_mlp(self, key, x):
        n, h, w, c = x.shape
        x = lif(x.permute(0, 3, 1, 2).flatten(2))
        x = self.bn(key + ".fc1_bn", self.conv(key + ".fc1_conv", x))
        x = self.bn(key + ".fc2_bn", self.conv(key + ".fc2_conv", lif(x)))
        return x.reshape(n, h, w, c)                                                       # transformer.py:829 (reinterpret)

    def pixel_decoder(self, feats):
        P, pd_m, pd = self.P, self.hd.pixel_decoder, "decode_head.pixel_decoder."
        sa = pd_m.encoder_cfg["layer_cfg"]["self_attn_cfg"]
        x = self.bn(pd + "encoder_in_proj.1", self.conv(pd + "encoder_in_proj.0", lif(feats[-1])))
        q = x.permute(0, 2, 3, 1)
        for l in range(pd_m.encoder_cfg["num_layers"]):
            k = pd + f"encoder.layers.{l}"
            q = q + P[k + ".gamma1"] * self.sepconv_spike(k + ".Conv", q, 3)
            q = q + P[k + ".gamma2"] * self.dcn(k + ".dcn", q, sa["group"], sa["dw_kernel_size"])
            q = q + P[k + ".gamma3"] * self.enc_mlp(k + ".ffn", q)
        memory = lif(q.permute(0, 3, 1, 2))
        y = self.bn(pd + "encoder_out_proj.1", self.conv(pd + "encoder_out_proj.0", memory))
        outs = [y]
        for i in range(len(feats) - 2, -1, -1):
            cur = self.bn(pd + f"lateral_convs.{i}.1", self.conv(pd + f"lateral_convs.{i}.0", lif(feats[i])))
            y = lif(cur + F.interpolate(y, size=cur.shape[-2:], mode="bilinear", align_corners=False))
            y = self.bn(pd + f"output_convs.{i}.1", self.conv(pd + f"output_convs.{i}.0", y, pad=1, groups=y.shape[1]))
            outs.append(y)
        return self.conv(pd + "mask_feature", lif(y)), memory, outs[:3]

    # -- transformer decoder + SDME (maskformer_head.py:498-586, detr_layers.py:491-559, mmcv_spike/transformer.py)
    def sine_pe(self, h, w, device):
        if (h, w) not in self.pe:
            nf = self.hd.positional_encoding_cfg["num_feats"]
            ones = torch.ones(1, h, w, dtype=torch.int)
            y_e, x_e = ones.cumsum(1, dtype=torch.float32), ones.cumsum(2, dtype=torch.float32)
            y_e = (y_e + 0.0) / (y_e[:, -1:, :] + 1e-6) * (2 * math.pi)
            x_e = (x_e + 0.0) / (x_e[:, :, -1:] + 1e-6) * (2 * math.pi)
            dim_t = torch.arange(nf, dtype=torch.float32)
            dim_t = 10000 ** (2 * (dim_t // 2) / nf)
            px, py = x_e[:, :, :, None] / dim_t, y_e[:, :, :, None] / dim_t
            px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).view(1, h, w, -1)
            py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).view(1, h, w, -1)
            self.pe[(h, w)] = torch.cat((py, px), dim=3).reshape(1, h * w, 2 * nf).to(device)
        return self.pe[(h, w)]

    def attn_block(self, key, query, kin, vin, heads, attn_mask=None):
        n, nq, dim = query.shape
        nk = kin.shape[1]

        def proj(t, nm):
            t = lif(t).permute(0, 2, 1)
            t = self.bn(key + f".{nm}_conv.1", self.conv(key + f".{nm}_conv.0", t))
            return lif(t.permute(0, 2, 1))

        d = dim // heads
        hs = lambda t: t.view(n, -1, heads, d).permute(0, 2, 1, 3)
        q, k, v = hs(proj(query, "q")), hs(proj(kin, "k")), hs(proj(vin, "v"))
        scores = torch.matmul(q, k.transpose(2, 3)) / (dim ** 0.5)
        if attn_mask is not None:
            scores = scores.masked_fill(attn_mask.reshape(n, heads, nq, nk), 0)
        out = torch.matmul(scores, v).permute(0, 2, 1, 3).reshape(n, nq, dim)
        out = lif(out).permute(0, 2, 1)
        return self.bn(key + ".out_conv.1", self.conv(key + ".out_conv.0", out)).permute(0, 2, 1)

    def dec_ffn(self, key, x):
        n, nq, c = x.shape
        o = lif(x).reshape(n, c, nq)                                                       # transformer.py:777 (reinterpret)
        o = self.bn(key + ".bn1", self.conv(key + ".fc1", o))
        o = self.bn(key + ".bn2", self.conv(key + ".fc2", lif(o)))
        return o.reshape(n, nq, c)                                                         # :781

    def head(self, feats):
        """-> (all_cls_scores [L,B,nq,K+1], all_mask_preds [L,B,nq,h,w])."""
        P, hd_m, hd = self.P, self.hd, "decode_head."
        T = self.bb.T
        mask_feat, memory, ms = self.pixel_decoder(feats)
        n = feats[0].shape[0]
        bs = n // T
        heads = hd_m.transformer_decoder_cfg["layer_cfg"]["self_attn_cfg"]["num_heads"]
        qf = P[hd + "query_feat.weight"].unsqueeze(0).repeat(n, 1, 1)
        qe = P[hd + "query_embed.weight"].unsqueeze(0).repeat(n, 1, 1)
        dec_in, dec_pe = [], []
        for i in range(3):
            dec_in.append(ms[i].flatten(2).permute(0, 2, 1) + P[hd + "level_embed.weight"][i].view(1, 1, -1))
            dec_pe.append(self.sine_pe(ms[i].shape[-2], ms[i].shape[-1], qf.device))
        outs = [qf]
        for i in range(hd_m.num_transformer_decoder_layers):
            key, kv, kpos = hd + f"transformer_decoder.layers.{i}", dec_in[i % 3], dec_pe[i % 3]
            qf = qf + self.attn_block(key + ".cross_attn.attn", qf + qe, kv + kpos, kv, heads)
            qf = qf + self.attn_block(key + ".self_attn.attn", qf + qe, qf + qe, qf, heads)
            qf = qf + self.dec_ffn(key + ".ffn", qf)
            outs.append(qf)
        od = torch.sigmoid(torch.stack(outs))
        L, _, nq, c = od.shape
        od_ = 4 * lif(od)
        cls = F.linear(od_, P[hd + "cls_embed.weight"], P[hd + "cls_embed.bias"]).view(L, T, bs, nq, -1).mean(1)
        m = 4 * lif(F.linear(od_, P[hd + "mask_embed.fc1.weight"]))
        m = 4 * lif(F.linear(m, P[hd + "mask_embed.fc2.weight"]))
        m = F.linear(m, P[hd + "mask_embed.fc_out.weight"], P[hd + "mask_embed.fc_out.bias"])
        sc = (4 * lif(od)).reshape(L * n, nq, c)
        sc = self.bn(hd + "shortcut_conv.1", self.conv(hd + "shortcut_conv.0", sc)).view(L, n, nq, c)
        m = (4 * lif(m + P[hd + "w"] * sc)).view(L, T, bs, nq, c)
        mf = mask_feat.view(T, bs, *mask_feat.shape[1:])
        return cls, torch.einsum("ltbqc,tbchw->ltbqhw", m, mf).mean(1)


# ------------------------------------------------------------------------------------------------ loss
LOSS = dict(cls_weight=1.0, bg_class_weight=0.1, mask_weight=20.0, gamma=2.0, alpha=0.25, dice_weight=1.0, dice_eps=1.0,
            cost_cls=1.0, cost_mask=20.0, cost_dice=1.0, cost_dice_eps=1.0, cost_focal_eps=1e-12)      # cfg:94-131
EPS32 = float(torch.finfo(torch.float32).eps)


def seg_to_instances(gt, ignore_index=255):
    """decode_heads/maskformer_head.py:77-105 for one label map [1,H,W] -> (labels [G], masks [G,H,W] long)."""
    labels = torch.unique(gt)
    labels = labels[labels != ignore_index]
    if labels.numel() == 0:
        return labels, torch.zeros((0,) + tuple(gt.shape[-2:]), dtype=torch.long, device=gt.device)
    return labels, (gt.reshape(1, *gt.shape[-2:]) == labels.view(-1, 1, 1)).long()


@torch.no_grad()
def batched_match(all_cls, all_masks, labels_list, masks_list, c=LOSS):
    """Hungarian matching of every (decoder output, image) pair with ONE device->host copy.
    -> (list over layers of list over images of (pos_inds, pos_gt_inds) on the device, avg_factor [L] on the device)."""
    from scipy.optimize import linear_sum_assignment

    L, B, nq = all_cls.shape[:3]
    hw = all_masks.shape[-2:]
    costs, shapes = [], []
    for i in range(B):
        gl, gm = labels_list[i], masks_list[i]
        G = int(gl.numel())
        shapes.append(G)
        if G == 0:
            continue
        g = F.interpolate(gm.unsqueeze(1).float(), hw, mode="nearest").flatten(1)           # maskformer_head.py:328-334
        p = all_masks[:, i].flatten(2).sigmoid()                                             # [L, nq, hw]
        n = p.shape[-1]
        cost_cls = -all_cls[:, i].softmax(-1)[:, :, gl] * c["cost_cls"]                      # match_cost.py:217-223
        neg = -(1 - p + c["cost_focal_eps"]).log() * (1 - c["alpha"]) * p.pow(c["gamma"])    # :271-293
        pos = -(p + c["cost_focal_eps"]).log() * c["alpha"] * (1 - p).pow(c["gamma"])
        cost_mask = (torch.einsum("lnc,mc->lnm", pos, g) + torch.einsum("lnc,mc->lnm", neg, 1 - g)) / n * c["cost_mask"]
        den = p.sum(-1)[:, :, None] + g.sum(-1)[None, None, :]                               # :360-370
        cost_dice = (1 - (2 * torch.einsum("lnc,mc->lnm", p, g) + c["cost_dice_eps"]) / (den + c["cost_dice_eps"])) * c["cost_dice"]
        costs.append(torch.stack([cost_cls, cost_mask, cost_dice]).sum(0).reshape(-1))       # hungarian_assigner.py:122-130
    if costs:
        host = torch.cat(costs).cpu()                                                        # the step's one D2H sync
    flat, off = [], 0
    counts = torch.zeros(L, dtype=torch.long)
    for i in range(B):
        G = shapes[i]
        for l in range(L):
            if G == 0:
                counts[l] += 1                                                               # max(num_pos, 1)
                flat.append((l, i, torch.zeros(0, dtype=torch.long), torch.zeros(0, dtype=torch.long)))
                continue
            cm = host[off + l * nq * G: off + (l + 1) * nq * G].view(nq, G)
            rows, cols = linear_sum_assignment(cm)
            rows, cols = torch.from_numpy(rows), torch.from_numpy(cols)
            order = torch.argsort(rows)                                                      # pos_inds = nonzero(...).unique(): sorted
            flat.append((l, i, rows[order], cols[order]))
            counts[l] += max(int(rows.numel()), 1)
        off += L * nq * G
    dev = all_cls.device
    packed = torch.cat([torch.stack([r, cidx]) for _, _, r, cidx in flat], 1).to(dev, non_blocking=True) if flat else None
    out = [[None] * B for _ in range(L)]
    o = 0
    for l, i, r, _ in flat:
        k = int(r.numel())
        out[l][i] = (packed[0, o:o + k], packed[1, o:o + k])
        o += k
    return out, counts.to(dev)


def losses_from_matching(all_cls, all_masks, labels_list, masks_list, match, num_total_masks, num_classes, c=LOSS):
    """_loss_by_feat_single (maskformer_head.py:410-496) for every decoder output, on the device."""
    L, B, nq = all_cls.shape[:3]
    dev = all_cls.device
    class_weight = torch.tensor([1.0] * num_classes + [c["bg_class_weight"]], device=dev)
    out = {}
    for l in range(L):
        labels = torch.full((B, nq), num_classes, dtype=torch.long, device=dev)
        sel_b, sel_q, targets = [], [], []
        for i in range(B):
            pos, gti = match[l][i]
            labels[i, pos] = labels_list[i][gti]
            sel_b.append(torch.full_like(pos, i)); sel_q.append(pos); targets.append(masks_list[i][gti])
        sel_b, sel_q, mask_targets = torch.cat(sel_b), torch.cat(sel_q), torch.cat(targets, 0)
        flat_labels = labels.flatten()
        ce = F.cross_entropy(all_cls[l].flatten(0, 1), flat_labels, weight=class_weight, reduction="none")
        loss_cls = c["cls_weight"] * ce.sum() / (class_weight[flat_labels].sum() + EPS32)
        ntm = num_total_masks[l].clamp(min=1)
        mp = all_masks[l][sel_b, sel_q]                                                      # == mask_preds[mask_weights > 0]
        if mask_targets.shape[0] == 0:
            loss_mask = loss_dice = mp.sum()
        else:
            mp = F.interpolate(mp.unsqueeze(1), mask_targets.shape[-2:], mode="bilinear", align_corners=False).squeeze(1)
            inp, tgt = mp.sigmoid().flatten(1), mask_targets.flatten(1).float()
            dice = 1 - (2 * (inp * tgt).sum(1) + c["dice_eps"]) / (inp.sum(1) + tgt.sum(1) + c["dice_eps"])
            loss_dice = c["dice_weight"] * dice.sum() / (ntm + EPS32)
            h, w = mp.shape[-2:]
            pred, target = mp.reshape(-1), mask_targets.reshape(-1).float()                  # "target is (1 - mask_targets)": class 0
            ps = pred.sigmoid()
            pt = (1 - ps) * target + ps * (1 - target)
            fw = (c["alpha"] * target + (1 - c["alpha"]) * (1 - target)) * pt.pow(c["gamma"])
            focal = F.binary_cross_entropy_with_logits(pred, target, reduction="none") * fw
            loss_mask = c["mask_weight"] * focal.sum() / (ntm * h * w + EPS32)
        pre = "" if l == L - 1 else f"d{l}."
        out[pre + "loss_cls"], out[pre + "loss_mask"], out[pre + "loss_dice"] = loss_cls, loss_mask, loss_dice
    return out


def head_losses(net, feats, gt_sem_seg, ignore_index=255):
    """mmseg MaskFormerHead.loss: forward of all 7 decoder outputs + loss_by_feat -> dict of 21 loss terms."""
    all_cls, all_masks = net.head(feats)
    return losses_from_outputs(net, all_cls, all_masks, gt_sem_seg, ignore_index)


def losses_from_outputs(net, all_cls, all_masks, gt_sem_seg, ignore_index=255):
    """loss_by_feat (maskformer_head.py:367-408) on top of the mmseg label conversion."""
    inst = [seg_to_instances(g, ignore_index) for g in gt_sem_seg]
    ll, ml = [a for a, _ in inst], [b for _, b in inst]
    match, counts = batched_match(all_cls.detach(), all_masks.detach(), ll, ml)
    ntm = counts.float()
    if torch.distributed.is_available() and torch.distributed.is_initialized():              # reduce_mean (dist_utils.py:59-65)
        torch.distributed.all_reduce(ntm.div_(torch.distributed.get_world_size()))
    return losses_from_matching(all_cls, all_masks, ll, ml, match, ntm, net.hd.num_classes)


# ------------------------------------------------------------------------------------------------ step
class GradBuckets:
    """Gradient all-reduce (mean) in ~25 MB buckets, each launched from an autograd hook as soon as the gradients of its
    parameters exist, so NCCL overlaps the rest of the backward pass.  Buckets follow reverse registration order, the
    approximate order in which the backward pass produces gradients (what torch DDP does, tools/dist_train.sh)."""

    def __init__(self, params, bucket_mb=25):
        import torch.distributed as dist

        self.dist, self.world = dist, dist.get_world_size()
        self.params = [p for p in params if p.requires_grad]
        self.buckets, cur, size = [], [], 0
        for p in reversed(self.params):
            cur.append(p); size += p.numel() * 4
            if size >= bucket_mb * (1 << 20):
                self.buckets.append(cur); cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.where = {id(p): bi for bi, b in enumerate(self.buckets) for p in b}
        self.flat = [torch.zeros(sum(p.numel() for p in b), device=b[0].device) for b in self.buckets]
        self.pending = [0] * len(self.buckets)
        self.work = []
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._hook)
        self.bytes = sum(f.numel() * 4 for f in self.flat)

    def start(self):
        self.pending = [len(b) for b in self.buckets]
        self.work = []

    def _hook(self, p):
        bi = self.where[id(p)]
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi):
        flat, o = self.flat[bi], 0
        for p in self.buckets[bi]:
            n = p.numel()
            if p.grad is None:
                flat[o:o + n].zero_()
            else:
                flat[o:o + n].copy_(p.grad.reshape(-1))
            o += n
        self.work.append((bi, self.dist.all_reduce(flat, async_op=True)))

    def finish(self):
        for bi, pend in enumerate(self.pending):            # parameters that received no gradient in this step
            if pend > 0:
                self.pending[bi] = 0
                self._launch(bi)
        for bi, w in self.work:
            w.wait()
            flat, o = self.flat[bi], 0
            flat.div_(self.world)
            for p in self.buckets[bi]:
                n = p.numel()
                if p.grad is not None:
                    p.grad.copy_(flat[o:o + n].view_as(p.grad))
                o += n


def param_groups(seg, lr=1e-3, weight_decay=0.005):
    """paramwise_cfg of the config (:140-154): backbone lr x 0.1; query_embed / query_feat / level_embed no decay."""
    groups = {}
    for name, p in seg.named_parameters():
        if not p.requires_grad:
            continue
        lr_m, wd_m = 1.0, 1.0
        if "backbone" in name:
            lr_m, wd_m = 0.1, 1.0
        for key in ("query_embed", "query_feat", "level_embed"):
            if key in name:
                lr_m, wd_m = 1.0, 0.0
        groups.setdefault((lr_m, wd_m), []).append(p)
    return [dict(params=ps, lr=lr * lm, weight_decay=weight_decay * wm) for (lm, wm), ps in groups.items()]


class _NetModule(torch.nn.Module):
    """The network part of the step (image -> 7 class-score / mask-logit maps) as a module over the segmentor's
    parameters, so that torch.cuda.make_graphed_callables can capture its forward and backward passes."""

    def __init__(self, seg, net):
        super().__init__()
        self.seg, self.net = seg, net

    def forward(self, img):
        return self.net.head(self.net.backbone(img))


class TrainStep:
    """One optimisation step of config 5: forward, 21 losses, backward, bucketed NCCL all-reduce, clip 0.01, AdamW.

    precision: "fp32" (exact library arithmetic: the parity setting), "tf32" (cuDNN / cuBLAS TF32 tensor cores) or
    "bf16" (autocast of the convolutions / matmuls; neurons, BatchNorm statistics and the loss stay fp32).
    graph: capture the network's forward and backward passes (~20 000 + ~30 000 small launches at batch 6, launch-bound
    when issued one by one) into two CUDA graphs per input shape; matching and the losses stay eager (their shapes
    depend on the labels).  With a graph the gradient buckets are launched when the backward graph has finished."""

    def __init__(self, seg, lr=1e-3, weight_decay=0.005, max_norm=0.01, precision="fp32", bucket_mb=25, graph=False):
        if next(seg.parameters()).device.type != "cuda":
            raise RuntimeError("TrainStep: move the model to a CUDA device first (no CPU path)")
        self.seg, self.max_norm, self.precision = seg, max_norm, precision
        self.net = Net(seg)
        self.graph, self._graphed = graph, {}
        self.opt = torch.optim.AdamW(param_groups(seg, lr, weight_decay), betas=(0.9, 0.999), fused=True)
        dist = torch.distributed
        self.buckets = GradBuckets(list(seg.parameters()), bucket_mb) if dist.is_available() and dist.is_initialized() \
            and dist.get_world_size() > 1 else None

    def losses(self, img, gt_sem_seg):
        """EncoderDecoder.loss (encoder_decoder.py:163-188): img fp32 [B,3,H,W], gt_sem_seg long [B,1,H,W] -> dict."""
        if not img.is_cuda:
            raise RuntimeError("TrainStep: CUDA tensors required (no CPU path)")
        tf32 = self.precision != "fp32"
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32, tf32
        try:
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.precision == "bf16", cache_enabled=not self.graph):
                if self.graph:
                    key = (tuple(img.shape), img.device.index)
                    if key not in self._graphed:
                        # capture runs the forward a few times: put the BatchNorm running statistics back afterwards
                        # (RepConv's border value reads them, sdtv2.py:68-74, so they are part of the forward's input)
                        saved = {k: b.clone() for k, b in self.seg.named_buffers()}
                        self._graphed[key] = torch.cuda.make_graphed_callables(_NetModule(self.seg, self.net), (img.clone(),),
                                                                               allow_unused_input=True)
                        with torch.no_grad():
                            for k, b in self.seg.named_buffers():
                                b.copy_(saved[k])
                    all_cls, all_masks = self._graphed[key](img)
                else:
                    all_cls, all_masks = self.net.head(self.net.backbone(img))
            return losses_from_outputs(self.net, all_cls.float(), all_masks.float(), gt_sem_seg, self.seg.decode_head.ignore_index)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old

    def __call__(self, img, gt_sem_seg, timing=None):
        """timing: optional dict that receives CUDA-event milliseconds of the phases of this step
        (forward_loss, backward, allreduce_exposed = time spent waiting for NCCL after the backward pass, optimizer)."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if timing is not None else None
        mark = (lambda i: ev[i].record()) if ev else (lambda i: None)
        self.opt.zero_grad(set_to_none=False)
        if self.buckets is not None:
            self.buckets.start()
        mark(0)
        losses = self.losses(img, gt_sem_seg)
        total = sum(v for k, v in losses.items() if "loss" in k)         # mmengine parse_losses
        mark(1)
        total.backward()
        mark(2)
        if self.buckets is not None:
            self.buckets.finish()
        mark(3)
        torch.nn.utils.clip_grad_norm_(self.seg.parameters(), self.max_norm, norm_type=2)     # cfg:153
        self.opt.step()
        mark(4)
        self.seg.invalidate()                                            # inference plans / graphs hold the old weights
        if ev:
            torch.cuda.synchronize()
            for name, i in (("forward_loss", 0), ("backward", 1), ("allreduce_exposed", 2), ("optimizer", 3)):
                timing[name] = timing.get(name, 0.0) + ev[i].elapsed_time(ev[i + 1])
        return losses
