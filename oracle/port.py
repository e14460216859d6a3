"""CPU restatement of the Spike2Former hot path (TEST INFRASTRUCTURE ONLY).

A functional, state_dict-driven rewrite of the reference's inference path in
plain fp32 PyTorch.  It exists so that the parity checker can travel to the GPU
box, where /root/reference does not exist.  It is pinned against the reference
itself: tests/test_oracle_port.py compares it with the reference files executed
by oracle/ref_loader.py (in the build container) and with the fixtures under
tests/golden/ (everywhere).

Tensors are kept as [T*B, ...] -- the reference's neurons have no recurrence
over T (one elementwise call on the whole [T,B,...] tensor), so
T is a batch dimension everywhere except the two `.mean(1)` reductions of the
head.  Every function cites the reference lines it follows.

`P` is a flat dict of tensors keyed like the reference checkpoint
("backbone.block3.0.attn.q_conv.0.body.0.weight", "decode_head.w", ...).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5


class Ctx:
    """Run-time knobs + optional taps.

    calibrate: batch-norms run in training mode with momentum 1 so that the
               running statistics in P become the batch statistics (the
               calibrated-random-init recipe of SURVEY.md section 8d).
    tap:       callable(name, pre_activation, spike_levels) called at every neuron.
    """

    def __init__(self, P, calibrate=False, tap=None, d_max=8.0, norm=8.0, mutate=None, train=False):
        self.P, self.calibrate, self.tap = P, calibrate, tap
        self.train = train         # training mode: batch-statistics BatchNorm (momentum 0.1) + surrogate-gradient neurons
        self.mutate = mutate       # callable(name, levels) -> levels: fault injection for the stability experiments
        self.d_max, self.norm = d_max, norm
        self.ties = 0
        self.neurons = 0
        self.elems = 0
        self.marks = None          # dict name -> tensor for non-neuron tensors, when set to {}

    def mark(self, name, t):
        if self.marks is not None:
            self.marks[name] = t
        return t


# ----------------------------------------------------------------------------- primitives
class QuantSTE(torch.autograd.Function):
    """`quant` (surrogate.py:522-538): forward round(clamp(i, 0, 8)); backward grad * 1[0 <= i <= 8]."""

    @staticmethod
    def forward(ctx, i, d_max):
        ctx.save_for_backward(i)
        ctx.d_max = d_max
        return torch.round(torch.clamp(i, min=0, max=d_max))

    @staticmethod
    def backward(ctx, grad_output):
        (i,) = ctx.saved_tensors
        g = grad_output.clone()
        g[i < 0] = 0
        g[i > ctx.d_max] = 0
        return g, None


def lif(cx: Ctx, name: str, x: torch.Tensor) -> torch.Tensor:
    """Q_IFNode(Quant()) after reset: neuron.py:459-460 (charge), :115-131 + surrogate.py:522-529
    (fire = round(clamp(v,0,8)), round half to even), :133-153 (soft reset, unused after), :197 (/8)."""
    v = 0.0 + x
    s = QuantSTE.apply(v, cx.d_max) if cx.train else torch.round(torch.clamp(v, min=0, max=cx.d_max))
    if cx.mutate is not None:
        s = cx.mutate(name, s)
    cx.neurons += 1
    cx.elems += x.numel()
    if cx.tap is not None:
        cx.tap(name, x.detach(), s.detach())
    return s / cx.norm


def bn(cx: Ctx, key: str, x: torch.Tensor) -> torch.Tensor:
    P = cx.P
    return F.batch_norm(x, P[key + ".running_mean"], P[key + ".running_var"], P[key + ".weight"], P[key + ".bias"],
                        training=cx.calibrate or cx.train, momentum=1.0 if cx.calibrate else 0.1, eps=BN_EPS)


def conv(cx: Ctx, key: str, x, stride=1, pad=0, groups=1):
    P = cx.P
    w = P[key + ".weight"]
    b = P.get(key + ".bias")
    if w.dim() == 3:
        return F.conv1d(x, w, b)
    return F.conv2d(x, w, b, stride=stride, padding=pad, groups=groups)


# ----------------------------------------------------------------------------- backbone
def rep_conv(cx, key, x):
    """RepConv.body + outer BN: sdtv2.py:111-132, BNAndPadLayer :48-89, outer BN :280-296."""
    P = cx.P
    y = conv(cx, key + ".0.body.0", x)
    k = key + ".0.body.1.bn"
    y = bn(cx, k, y)
    pad_val = P[k + ".bias"].detach() - P[k + ".running_mean"] * P[k + ".weight"].detach() / \
        torch.sqrt(P[k + ".running_var"] + BN_EPS)                                    # sdtv2.py:68-74 (detached)
    y = F.pad(y, [1, 1, 1, 1])
    pv = pad_val.view(1, -1, 1, 1)
    y[:, :, 0:1, :] = pv
    y[:, :, -1:, :] = pv
    y[:, :, :, 0:1] = pv
    y[:, :, :, -1:] = pv
    y = conv(cx, key + ".0.body.2.0", y, groups=y.shape[1])
    y = conv(cx, key + ".0.body.2.1", y)
    y = bn(cx, key + ".0.body.2.2", y)
    return bn(cx, key + ".1", y)


def downsample(cx, key, x, stride, pad, first):
    """MS_DownSampling.forward: sdtv2.py:412-421."""
    if not first:
        x = lif(cx, key + ".encode_spike", x)
    return bn(cx, key + ".encode_bn", conv(cx, key + ".encode_conv", x, stride, pad))


def sep_conv(cx, key, x):
    """SepConv.forward: sdtv2.py:167-180."""
    x = lif(cx, key + ".spike1", x)
    x = bn(cx, key + ".bn1", conv(cx, key + ".pwconv1", x))
    x = lif(cx, key + ".spike2", x)
    x = conv(cx, key + ".dwconv", x, pad=3, groups=x.shape[1])
    return bn(cx, key + ".bn2", conv(cx, key + ".pwconv2", x))


def conv_block(cx, key, x):
    """MS_ConvBlock.forward: sdtv2.py:207-219."""
    x = sep_conv(cx, key + ".Conv", x) + x
    feat = x
    x = lif(cx, key + ".spike1", x)
    x = bn(cx, key + ".bn1", conv(cx, key + ".conv1", x, pad=1))
    x = lif(cx, key + ".spike2", x)
    x = bn(cx, key + ".bn2", conv(cx, key + ".conv2", x, pad=1))
    return feat + x


def sdsa(cx, key, x, heads):
    """MS_Attention_RepConv_qkv_id.forward: sdtv2.py:298-344."""
    n, c, h, w = x.shape
    tok = h * w
    d = c // heads
    scale = d ** -0.5
    x = lif(cx, key + ".head_spike", x)
    q = rep_conv(cx, key + ".q_conv", x)
    k = rep_conv(cx, key + ".k_conv", x)
    v = rep_conv(cx, key + ".v_conv", x)

    def split(t, nm):
        t = lif(cx, key + nm, t).flatten(2)                      # [n, C, N]
        return t.transpose(-1, -2).reshape(n, tok, heads, d).permute(0, 2, 1, 3).contiguous()

    q, k, v = split(q, ".q_spike"), split(k, ".k_spike"), split(v, ".v_spike")
    kv = k.transpose(-2, -1) @ v                                 # sdtv2.py:335
    y = (q @ kv) * scale                                         # :336
    y = y.transpose(2, 3).reshape(n, c, tok).contiguous()
    y = lif(cx, key + ".attn_spike", y)
    return rep_conv(cx, key + ".proj_conv", y.reshape(n, c, h, w))


def ms_mlp(cx, key, x):
    """MS_MLP.forward (backbone): sdtv2.py:242-255."""
    n, c, h, w = x.shape
    x = lif(cx, key + ".fc1_spike", x.flatten(2))
    x = bn(cx, key + ".fc1_bn", conv(cx, key + ".fc1_conv", x)).contiguous()
    x = lif(cx, key + ".fc2_spike", x)
    x = bn(cx, key + ".fc2_bn", conv(cx, key + ".fc2_conv", x))
    return x.reshape(n, c, h, w).contiguous()


def ms_block(cx, key, x, heads):
    """MS_Block.forward: sdtv2.py:377-383."""
    x = x + sdsa(cx, key + ".attn", x, heads)
    return x + ms_mlp(cx, key + ".mlp", x)


def backbone_forward(cx, cfg, img):
    """Spiking_vit_MetaFormer.forward_features, decode_mode 'Qsnn': sdtv2.py:614-651.
    Returns four [T*B, C, H, W] maps (the reference returns them as [T,B,C,H,W])."""
    T = cfg["T"]
    heads = cfg["num_heads"]
    x = img.unsqueeze(0).repeat(T, 1, 1, 1, 1).flatten(0, 1)
    b = "backbone."
    x = downsample(cx, b + "downsample1_1", x, 2, 3, True)
    x = conv_block(cx, b + "ConvBlock1_1.0", x); x1 = x
    x = downsample(cx, b + "downsample1_2", x, 2, 1, False)
    x = conv_block(cx, b + "ConvBlock1_2.0", x); x2 = x
    x = downsample(cx, b + "downsample2", x, 2, 1, False)
    x = conv_block(cx, b + "ConvBlock2_1.0", x)
    x = conv_block(cx, b + "ConvBlock2_2.0", x); x3 = x
    x = downsample(cx, b + "downsample3", x, 2, 1, False)
    for j in range(6):
        x = ms_block(cx, b + f"block3.{j}", x, heads)
    x = downsample(cx, b + "downsample4", x, 1, 1, False)
    for j in range(2):
        x = ms_block(cx, b + f"block4.{j}", x, heads)
    return [x1, x2, x3, x]


# ----------------------------------------------------------------------------- pixel decoder
def sepconv_spike(cx, key, x, k):
    """SepConv_Spike.forward: mmcv_spike/SNN_core.py:47-63.  x: [n,H,W,C] channels-last -> same."""
    x = x.permute(0, 3, 1, 2).contiguous()
    x = lif(cx, key + ".spike1", x)
    x = bn(cx, key + ".pwconv1.1", conv(cx, key + ".pwconv1.0", x))
    x = lif(cx, key + ".spike2", x)
    x = bn(cx, key + ".dwconv.1", conv(cx, key + ".dwconv.0", x, pad=(k - 1) // 2, groups=x.shape[1]))
    x = lif(cx, key + ".spike3", x)
    x = bn(cx, key + ".pwconv2.1", conv(cx, key + ".pwconv2.0", x))
    return x.permute(0, 2, 3, 1).contiguous()


def dcnv3_core(x, offset, mask, ksz, stride, pad, dil, group, gch, offset_scale):
    """dcnv3_core_pytorch: ops_dcnv3/functions/dcnv3_func.py:147-189 with _get_reference_points
    :91-119 and _generate_dilation_grids :122-144.  x [n,H,W,C]; offset [n,H,W,G*K*2]; mask [n,H,W,G*K]."""
    x = F.pad(x, [0, 0, pad, pad, pad, pad])
    n, hin, win, _ = x.shape
    _, hout, wout, _ = offset.shape
    half = (dil * (ksz - 1)) // 2
    ry = torch.linspace(half + 0.5, half + 0.5 + (hout - 1) * stride, hout, dtype=torch.float32)
    rx = torch.linspace(half + 0.5, half + 0.5 + (wout - 1) * stride, wout, dtype=torch.float32)
    ref_y, ref_x = torch.meshgrid(ry, rx)
    ref_y = ref_y.reshape(-1)[None] / hin
    ref_x = ref_x.reshape(-1)[None] / win
    ref = torch.stack((ref_x, ref_y), -1).reshape(1, hout, wout, 1, 2)
    lin = torch.linspace(-half, -half + (ksz - 1) * dil, ksz, dtype=torch.float32)
    gx, gy = torch.meshgrid(lin, lin)
    grid = torch.stack([gx / win, gy / hin], -1).reshape(-1, 1, 2).repeat(1, group, 1).permute(1, 0, 2)
    grid = grid.reshape(1, 1, 1, group * ksz * ksz, 2)
    spatial_norm = torch.tensor([win, hin]).reshape(1, 1, 1, 2).repeat(1, 1, 1, group * ksz * ksz)
    loc = (ref + grid * offset_scale).repeat(n, 1, 1, 1, 1).flatten(3, 4) + offset * offset_scale / spatial_norm
    pts = ksz * ksz
    sgrid = 2 * loc - 1
    xin = x.view(n, hin * win, group * gch).transpose(1, 2).reshape(n * group, gch, hin, win)
    sg = sgrid.view(n, hout * wout, group, pts, 2).transpose(1, 2).flatten(0, 1)
    samp = F.grid_sample(xin, sg, mode="bilinear", padding_mode="zeros", align_corners=False)
    m = mask.view(n, hout * wout, group, pts).transpose(1, 2).reshape(n * group, 1, hout * wout, pts)
    out = (samp * m).sum(-1).view(n, group * gch, hout * wout)
    return out.transpose(1, 2).reshape(n, hout, wout, -1).contiguous()


def dcn(cx, key, inp, group, dw_k, ksz=3):
    """DCNv3_pytorch.forward: ops_dcnv3/modules/dcnv3.py:198-233.  inp [n,H,W,C]."""
    n, h, w, c = inp.shape
    x = sepconv_spike(cx, key + ".input_proj", inp, dw_k)
    x1 = inp.permute(0, 3, 1, 2).contiguous()
    x1 = lif(cx, key + ".dw_spike", x1)
    x1 = bn(cx, key + ".dw_conv.1", conv(cx, key + ".dw_conv.0", x1, pad=(dw_k - 1) // 2, groups=c))
    x1 = lif(cx, key + ".offset_spike", x1)
    # NCHW conv outputs *reinterpreted* (not permuted) as [n,H,W,-1]: dcnv3.py:214-215
    offset = bn(cx, key + ".offset.1", conv(cx, key + ".offset.0", x1)).reshape(n, h, w, -1)
    mask = bn(cx, key + ".mask.1", conv(cx, key + ".mask.0", x1)).reshape(n, h, w, group, -1).reshape(n, h, w, -1)
    mask = lif(cx, key + ".mask_spike", mask)
    cx.mark(key + ".x", x); cx.mark(key + ".offset", offset)
    x = dcnv3_core(x, offset, mask, ksz, 1, 1, 1, group, c // group, 1.0)
    return sepconv_spike(cx, key + ".output_proj", x, dw_k)


def enc_mlp(cx, key, x):
    """MS_MLP.forward (head): mmcv_spike/transformer.py:817-831.  x [n,H,W,C]; output reinterpreted (:829)."""
    n, h, w, c = x.shape
    x = x.permute(0, 3, 1, 2).contiguous().flatten(2)
    x = lif(cx, key + ".fc1_spike", x)
    x = bn(cx, key + ".fc1_bn", conv(cx, key + ".fc1_conv", x))
    x = lif(cx, key + ".fc2_spike", x)
    x = bn(cx, key + ".fc2_bn", conv(cx, key + ".fc2_conv", x))
    return x.reshape(n, h, w, c)


def pixel_decoder_forward(cx, cfg, feats):
    """DCNTransformerEncoderPixelDecoder.forward: pixel_decoder.py:417-472;
    encoder layer: detr_layers.py:334-339."""
    P = cx.P
    pd = "decode_head.pixel_decoder."
    enc = cfg["encoder"]
    sa = enc["layer_cfg"]["self_attn_cfg"]
    x = lif(cx, pd + "last_feat_conv_spike", feats[-1])
    x = bn(cx, pd + "encoder_in_proj.1", conv(cx, pd + "encoder_in_proj.0", x))
    q = x.permute(0, 2, 3, 1)
    for l in range(enc["num_layers"]):
        k = pd + f"encoder.layers.{l}"
        q = q + P[k + ".gamma1"] * sepconv_spike(cx, k + ".Conv", q, 3)
        q = q + P[k + ".gamma2"] * dcn(cx, k + ".dcn", q, sa["group"], sa["dw_kernel_size"])
        q = q + P[k + ".gamma3"] * enc_mlp(cx, k + ".ffn", q)
    memory = q.permute(0, 3, 1, 2).contiguous()
    memory = lif(cx, pd + "encoder_out_proj_spike", memory)
    y = bn(cx, pd + "encoder_out_proj.1", conv(cx, pd + "encoder_out_proj.0", memory))
    cx.mark(pd + "y0", y)
    outs = [y]
    for i in range(len(feats) - 2, -1, -1):
        x = lif(cx, pd + f"lateral_convs_spike.{i}", feats[i])
        cur = bn(cx, pd + f"lateral_convs.{i}.1", conv(cx, pd + f"lateral_convs.{i}.0", x))
        y = cur + F.interpolate(y, size=cur.shape[-2:], mode="bilinear", align_corners=False)
        y = lif(cx, pd + f"output_convs_spike.{i}", y)
        y = bn(cx, pd + f"output_convs.{i}.1", conv(cx, pd + f"output_convs.{i}.0", y, pad=1, groups=y.shape[1]))
        outs.append(y)
        cx.mark(pd + f"y{len(outs) - 1}", y)
    y = lif(cx, pd + "mask_feature_spike", y)
    mask_feature = cx.mark(pd + "mask_feature", conv(cx, pd + "mask_feature", y))
    return mask_feature, memory, outs[:3]


# ----------------------------------------------------------------------------- transformer decoder
def sine_pe(num_feats, h, w, temperature=10000, scale=2 * math.pi, eps=1e-6):
    """SinePositionalEncoding.forward with an all-valid mask, normalize=True:
    positional_encoding.py:60-104.  Returns [1, 2*num_feats, h, w]."""
    not_mask = torch.ones(1, h, w, dtype=torch.int)
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    y_embed = (y_embed + 0.0) / (y_embed[:, -1:, :] + eps) * scale
    x_embed = (x_embed + 0.0) / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).view(1, h, w, -1)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).view(1, h, w, -1)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def attn_block(cx, key, query, kin, vin, heads, attn_mask=None):
    """{Cross,}MultiHeadAttentionBlock.forward: mmcv_spike/transformer.py:237-278 / 318-361.
    query [n,nq,C]; kin, vin [n,nk,C].  No softmax; scores / sqrt(embed_dim)."""
    n, nq, dim = query.shape
    nk = kin.shape[1]

    def proj(t, nm):
        t = lif(cx, key + f".{nm}_conv_spike", t).permute(0, 2, 1)
        t = bn(cx, key + f".{nm}_conv.1", conv(cx, key + f".{nm}_conv.0", t))
        return lif(cx, key + f".{nm}_spike", t.permute(0, 2, 1).contiguous())

    q, k, v = proj(query, "q"), proj(kin, "k"), proj(vin, "v")
    d = dim // heads
    hs = lambda t: torch.stack(torch.split(t, d, dim=2), dim=2).permute(0, 2, 1, 3).contiguous()
    qs, ks, vs = hs(q), hs(k), hs(v)
    scores = torch.matmul(qs, ks.transpose(2, 3)) / (dim ** 0.5)
    if attn_mask is not None:
        scores = scores.masked_fill(attn_mask.reshape(n, heads, nq, nk), 0)
    out = torch.matmul(scores, vs)
    out = torch.cat(torch.split(out, 1, dim=1), dim=3).squeeze(1)
    out = lif(cx, key + ".attn_spike", out).permute(0, 2, 1).contiguous()
    out = bn(cx, key + ".out_conv.1", conv(cx, key + ".out_conv.0", out))
    return out.permute(0, 2, 1).contiguous()


def dec_ffn(cx, key, x):
    """MSDA_FFN.forward: mmcv_spike/transformer.py:768-784 (both reshapes reinterpret memory)."""
    n, nq, c = x.shape
    o = lif(cx, key + ".fc1_spike", x).reshape(n, c, nq)
    o = bn(cx, key + ".bn1", conv(cx, key + ".fc1", o))
    o = lif(cx, key + ".fc2_spike", o)
    o = bn(cx, key + ".bn2", conv(cx, key + ".fc2", o))
    return o.reshape(n, nq, c)


def decoder_layer(cx, key, query, kv, qpos, kpos, heads, cross_mask=None):
    """DetrTransformerDecoderLayer.forward: detr_layers.py:491-559 and the
    MultiheadAttention wrapper mmcv_spike/transformer.py:561-638."""
    ca = attn_block(cx, key + ".cross_attn.attn", query + qpos, kv + kpos, kv, heads, cross_mask)
    query = query + ca
    sa = attn_block(cx, key + ".self_attn.attn", query + qpos, query + qpos, query, heads)
    query = query + sa
    return cx.mark(key + ".out", query + dec_ffn(cx, key + ".ffn", query))


def head_forward(cx, cfg, feats, T, all_layers=True):
    """mmdet MaskFormerHead.forward: dense_heads/maskformer_head.py:498-586.
    feats: list of [T*B, C, H, W].  Returns (all_cls_scores [L,B,nq,K+1], all_mask_preds [L,B,nq,h,w])."""
    P = cx.P
    hd = "decode_head."
    mask_feat, memory, ms = pixel_decoder_forward(cx, cfg["pixel_decoder"], feats)
    n = feats[0].shape[0]
    bs = n // T
    heads = cfg["transformer_decoder"]["layer_cfg"]["self_attn_cfg"]["num_heads"]
    nfeat = cfg["positional_encoding"]["num_feats"]
    qf = P[hd + "query_feat.weight"].unsqueeze(0).repeat(n, 1, 1)
    qe = P[hd + "query_embed.weight"].unsqueeze(0).repeat(n, 1, 1)
    dec_in, dec_pe = [], []
    for i in range(3):
        di = ms[i].flatten(2).permute(0, 2, 1) + P[hd + "level_embed.weight"][i].view(1, 1, -1)
        pe = sine_pe(nfeat, ms[i].shape[-2], ms[i].shape[-1]).flatten(2).permute(0, 2, 1)
        dec_in.append(di)
        dec_pe.append(pe)
    outs = [qf]
    nl = cfg["transformer_decoder"]["num_layers"]
    for i in range(nl):
        lvl = i % 3
        qf = decoder_layer(cx, hd + f"transformer_decoder.layers.{i}", qf, dec_in[lvl], qe, dec_pe[lvl], heads)
        outs.append(qf)
    od = torch.stack(outs)                                           # [L, n, nq, C]
    L, _, nq, c = od.shape
    # ---- SDME: maskformer_head.py:571-582
    od = torch.sigmoid(od)
    od_ = 4 * lif(cx, hd + "decoder_out_spike", od)
    cls = F.linear(od_, P[hd + "cls_embed.weight"], P[hd + "cls_embed.bias"]).view(L, T, bs, nq, -1).mean(1)
    m = F.linear(od_, P[hd + "mask_embed.fc1.weight"])                # SNN_core.py:116-123
    m = lif(cx, hd + "mask_embed.spike1", m) * 4
    m = F.linear(m, P[hd + "mask_embed.fc2.weight"])
    m = lif(cx, hd + "mask_embed.spike2", m) * 4
    m = F.linear(m, P[hd + "mask_embed.fc_out.weight"], P[hd + "mask_embed.fc_out.bias"])
    sc = (4 * lif(cx, hd + "shortcut_conv_spike", od)).reshape(L * n, nq, c)
    sc = bn(cx, hd + "shortcut_conv.1", conv(cx, hd + "shortcut_conv.0", sc)).view(L, n, nq, c).contiguous()
    m = m + P[hd + "w"] * sc
    m = 4 * lif(cx, hd + "mask_embed_spike", m)
    m = m.view(L, T, bs, nq, c)
    mf = mask_feat.view(T, bs, *mask_feat.shape[1:])
    masks = torch.einsum("ltbqc,tbchw->ltbqhw", m, mf).mean(1)
    return cls, masks


def predict(cx, cfg, img):
    """EncoderDecoder.encode_decode -> mmseg MaskFormerHead.predict:
    encoder_decoder.py:125-133, decode_heads/maskformer_head.py:138-180.  Returns [B,K,H,W] seg logits."""
    T = cfg["backbone"]["T"]
    feats = backbone_forward(cx, cfg["backbone"], img)
    cls, masks = head_forward(cx, cfg["decode_head"], feats, T)
    cls, mp = cx.mark("decode_head.cls_score", cls[-1]), cx.mark("decode_head.mask_pred", masks[-1])
    mp = F.interpolate(mp, size=img.shape[-2:], mode="bilinear", align_corners=False)
    score = F.softmax(cls, dim=-1)[..., :-1]
    return torch.einsum("bqc,bqhw->bchw", score, mp.sigmoid())


# ----------------------------------------------------------------------------- kernel-level oracles
def nilif_reference(x, scale=None, shift=None, residual=None, v0=None, d_max=8.0, T=1):
    """Multi-step generalisation (the north star's kernel (a)): x [T, N(, C)];
    per step v += affine(x_t) [+ residual]; s = rne(clamp(v,0,D)); v -= s.  Returns (levels int8 [T,...], v_final).
    T=1, v0=0 is exactly lif() above."""
    v = torch.zeros_like(x[0]) if v0 is None else v0.clone()
    out = []
    for t in range(T):
        u = x[t]
        if scale is not None:
            u = u * scale + shift
        if residual is not None:
            u = u + residual[t]
        v = v + u
        s = torch.round(torch.clamp(v, 0, d_max))
        v = v - s * 1.0
        out.append(s.to(torch.int8))
    return torch.stack(out), v


def count_ties(x, d_max=8.0):
    """Pre-activations that sit exactly on a rounding tie inside the clamp range."""
    fr = x - torch.floor(x)
    return int(((fr == 0.5) & (x > 0) & (x < d_max)).sum())


# --------------------------------------------------------------------------- data preprocessor
def data_preprocess(inputs, mean=None, std=None, bgr_to_rgb=False, rgb_to_bgr=False, size=None, size_divisor=None,
                    pad_val=0):
    """SegDataPreProcessor.forward (mmseg/models/data_preprocessor.py:109-152) for a list of uint8 [3,H,W] images:
    channel flip (:121-122), .float() (:124), (x - mean) / std (:125-126), then stack_batch's right/bottom padding
    with pad_val (mmseg/utils/misc.py:68-93) -> fp32 [B,3,Hp,Wp]."""
    if bgr_to_rgb or rgb_to_bgr:
        inputs = [i[[2, 1, 0], ...] for i in inputs]
    inputs = [i.float() for i in inputs]
    if mean is not None:
        m = torch.tensor(mean).view(-1, 1, 1)
        sd = torch.tensor(std).view(-1, 1, 1)
        inputs = [(i - m) / sd for i in inputs]
    hmax = max(i.shape[-2] for i in inputs)
    wmax = max(i.shape[-1] for i in inputs)
    if size_divisor is not None and size_divisor > 1:
        hmax = (hmax + size_divisor - 1) // size_divisor * size_divisor
        wmax = (wmax + size_divisor - 1) // size_divisor * size_divisor
    out = []
    for t in inputs:
        if size is not None:
            pw, ph = max(size[-1] - t.shape[-1], 0), max(size[-2] - t.shape[-2], 0)
        elif size_divisor is not None:
            pw, ph = max(wmax - t.shape[-1], 0), max(hmax - t.shape[-2], 0)
        else:
            pw = ph = 0
        out.append(F.pad(t, (0, pw, 0, ph), value=pad_val))
    return torch.stack(out, dim=0)
