"""Execution engine: runs the Spike2Former forward on the sm_100a kernels (no torch compute ops on the path).

Data layout in HBM (n = T*B images, channels-last everywhere):
  * spikes      int8 levels 0..8          [n, H, W, C]   (1 B / neuron; reference value = level / 8)
  * streams     fp32 residual streams     [n, H, W, C]
  * tokens      the same buffers viewed   [n, H*W, C]
Every kernel that produces a residual stream also emits the NI-LIF levels of that stream from its
epilogue, because every consumer of a stream in the reference starts with a Q_IFNode.

A `Plan` holds the folded / re-laid-out parameters of one model on one device.  Forward functions take
an optional `probe` (tests only): it sees every neuron's spikes (and pre-activations) and may replace
them, which is how the teacher-forced parity tests drive each unit with the oracle's tensors.
"""
from __future__ import annotations

import contextlib
import math

import torch

from . import fold, ops
from .ops import NORM

INV = 1.0 / NORM
USE_TC = True            # spike-operand layers run on the tcgen05 kernel (the CUDA-core kernel covers the rest)
import os as _os
TC_PIECES = int(_os.environ.get("S2F_TC_PIECES", "3"))   # int8 digit planes per weight: 3 = 21-bit fixed point (default), 2 = 14-bit fast mode
FUSE_FPN = _os.environ.get("S2F_FUSE_FPN", "1") != "0"           # finest FPN merges on the fp16 kernel (csrc/fpn_tc.cu)
FUSE_SEPCONV = _os.environ.get("S2F_FUSE_SEPCONV", "1") != "0"   # SepConv dw7x7 + pwconv2 as one launch (csrc/sepconv_tc.cu)
TC_MIN_ROWS = 1          # every eligible layer runs on the tensor-core kernel whatever the batch: its integer accumulation
                         # is order-independent, so an image's result never depends on how many images share the launch


# ------------------------------------------------------------------------------------------------ plan
class Gemm:
    """A conv / linear layer prepared for the kernels: weights [Cout, K] (Cin fastest) + folded affine."""

    def __init__(self, w2d, scale, shift, cin, k=1, stride=1, pad=0, device="cuda", spike_in=True):
        self.cout, self.cin, self.k, self.stride, self.pad = int(w2d.shape[0]), int(cin), k, stride, pad
        self.w = ops.pad_rows4(fold.f32(w2d, device))
        self.scale, self.shift = fold.f32(scale, device), fold.f32(shift, device)
        # MACs per output row the REFERENCE's layer(s) perform for what this launch computes (bench.py's roofline counts
        # algorithmic work: a re-parameterised RepConv executes 9*C*C here but is 2*C*C + 9*C there; zero-padded
        # channels are not work either).  None: the executed count is the algorithmic one.
        self.alg_macs = None
        # tensor-core form: int8 digit planes + scale with the packer's power-of-two row scale folded in
        self.tc = None
        if spike_in and USE_TC and ops.tc_eligible(self.cin, k, stride) and pad == (k - 1) // 2:
            packed, rowscale = ops.pack_weights_i8(w2d.to(torch.float32), k * k, self.cin, TC_PIECES)
            self.tc = (packed.to(device), fold.f32(scale.double().cpu() * rowscale.double(), device))

    def __call__(self, a, n, H, W, residual=None, f32=False, spike=False, transposed=False, a_scale=INV, up_prev=None, **kw):
        if self.tc is not None and a.dtype == torch.int8 and not kw and n * H * W >= TC_MIN_ROWS:
            sc = self.tc[1] if a_scale == 1.0 else self._scaled(a_scale)
            return ops.gemm_tc(a, self.tc[0], n=n, H=H, W=W, Cin=self.cin, Cout=self.cout, scale=sc, shift=self.shift,
                               k=self.k, stride=self.stride, pad=self.pad, pieces=TC_PIECES, residual=residual,
                               want_f32=f32, want_spike=spike, transposed=transposed, up_prev=up_prev,
                               alg_macs=self.alg_macs)
        if up_prev is not None:
            raise RuntimeError("fused FPN merge needs the tensor-core path")
        return self._simt(a, n, H, W, residual, f32, spike, transposed, a_scale, **kw)

    def _scaled(self, a_scale):
        cache = self.__dict__.setdefault("_sc", {})
        if a_scale not in cache:
            cache[a_scale] = (self.tc[1] * a_scale).contiguous()
        return cache[a_scale]

    def _simt(self, a, n, H, W, residual=None, f32=False, spike=False, transposed=False, a_scale=INV, **kw):
        return ops.conv_simt(a, self.w, n=n, H=H, W=W, Cin=self.cin, Cout=self.cout, k=self.k, stride=self.stride,
                             pad=self.pad, scale=self.scale, shift=self.shift, residual=residual, a_scale=a_scale,
                             want_f32=f32, want_spike=spike, transposed=transposed, **kw)


class Dw:
    def __init__(self, w, scale, shift, device):
        self.k, self.c = int(w.shape[-1]), int(w.shape[0])
        self.w = fold.f32(fold.dw_taps(w), device)
        self.scale = fold.f32(scale, device) if scale is not None else None
        self.shift = fold.f32(shift, device) if shift is not None else None

    def __call__(self, a, n, H, W, f32=False, spike=False):
        return ops.dwconv(a, self.w, n=n, H=H, W=W, C_=self.c, k=self.k, scale=self.scale, shift=self.shift,
                          want_f32=f32, want_spike=spike)


class SepDwPw:
    """SepConv's `dwconv` + `pwconv2` + `bn2` (sdtv2.py:176-179) as one launch (csrc/sepconv_tc.cu): the real-valued
    stencil output stays in shared memory as fp16 hi + lo and the 1x1 runs on tcgen05 kind::f16."""

    def __init__(self, sd, dw_key, pw_key, bn_key, device):
        wd = sd[dw_key + ".weight"]
        wp = fold.w_khwc(sd[pw_key + ".weight"])
        self.k, self.cm, self.cout = int(wd.shape[-1]), int(wd.shape[0]), int(wp.shape[0])
        s, t = fold.conv_bn(sd, pw_key, bn_key)
        packed, rowscale = ops.pack_pw_f16(wp)
        # |stencil output| <= sum |w_dw| (levels / 8 <= 1): the power of two that keeps the fp16 hi part below 2^15
        bound = float(wd.double().abs().sum(dim=(1, 2, 3)).max().clamp_min(1e-30))
        self.a_pre = float(2.0 ** max(-8, min(8, math.floor(math.log2(32000.0 / bound)))))
        self.w_dw = fold.f32(fold.dw_taps(wd), device)
        self.packed = packed.to(device)
        self.scale = fold.f32(s * rowscale.double() / self.a_pre, device)
        self.shift = fold.f32(t, device)

    @staticmethod
    def eligible(sd, dw_key, pw_key):
        cm, cout = int(sd[dw_key + ".weight"].shape[0]), int(sd[pw_key + ".weight"].shape[0])
        return USE_TC and FUSE_SEPCONV and cm % 64 == 0 and 16 <= cout <= 128 and cout % 4 == 0 and sd.get(dw_key + ".bias") is None

    def __call__(self, a, n, H, W, residual=None, f32=True, spike=True):
        return ops.sepconv_dwpw(a, self.w_dw, self.packed, n=n, H=H, W=W, Cm=self.cm, Cout=self.cout, k=self.k,
                                scale=self.scale, shift=self.shift, a_pre=self.a_pre, residual=residual,
                                want_f32=f32, want_spike=spike)


class FpnMergeF16:
    """lateral 1x1 + BN + bilinear x2 + add + NI-LIF of the two finest FPN levels (pixel_decoder.py:451-462) on
    csrc/fpn_tc.cu: tcgen05 kind::f16 with fp16 hi / lo weights, one fp32 accumulator per output."""

    def __init__(self, gm, device):
        w = gm.w[:, :gm.cin].detach().double().cpu()
        w64 = torch.zeros(gm.cout, 64, dtype=torch.float64)
        w64[:, :gm.cin] = w
        packed, rowscale = ops.pack_pw_f16(w64)
        self.cin, self.cout = gm.cin, gm.cout
        self.packed = packed.to(device)
        self.scale = fold.f32(gm.scale.double().cpu() * rowscale.double() * INV, device)
        self.shift = gm.shift
        self.alg_macs = gm.alg_macs

    @staticmethod
    def eligible(gm):
        return USE_TC and FUSE_FPN and gm.k == 1 and gm.cout == 256 and gm.cin % 16 == 0 and gm.cin <= 64

    def __call__(self, a, prev, n, H, W):
        return ops.fpn_merge_f16(a, self.packed, prev, n=n, H=H, W=W, Cin=self.cin, Cout=self.cout, scale=self.scale,
                                 shift=self.shift, alg_macs=self.alg_macs)


class StemU8:
    """The first MS_DownSampling (7x7 / 2, no LIF in front: sdtv2.py:412-421) with SegDataPreProcessor folded in
    (data_preprocessor.py:121-126): weights W / std as int8 digit planes over the raw pixel bytes, and the mean
    subtraction as a shift tabulated for every pattern of kernel rows / columns cut off by the image border
    (the reference zero-pads the *normalised* image).  See csrc/stem_u8.cu."""

    def __init__(self, sd, mean, std, swap_rb, chw, device):
        w = sd["downsample1_1.encode_conv.weight"].double()                 # [Cout, 3, 7, 7]
        s, t = fold.conv_bn(sd, "downsample1_1.encode_conv", "downsample1_1.encode_bn")
        cout = int(w.shape[0])
        mean = torch.tensor([float(torch.tensor(float(v), dtype=torch.float32)) for v in mean], dtype=torch.float64)
        std = torch.tensor([float(torch.tensor(float(v), dtype=torch.float32)) for v in std], dtype=torch.float64)
        # stored channel cs of the image is the model's channel cm = 2 - cs after inputs[[2, 1, 0]]
        cm_of = [2, 1, 0] if swap_rb else [0, 1, 2]
        w2d = torch.zeros(cout, 192, dtype=torch.float64)
        mean_k = torch.zeros(192, dtype=torch.float64)
        kh_of = torch.full((192,), -1, dtype=torch.long)
        kw_of = torch.full((192,), -1, dtype=torch.long)
        for kh in range(7):
            for j in range(21):
                kw, cs = ((j % 7), (j // 7)) if chw else ((j // 3), (j % 3))
                cm = cm_of[cs]
                k = kh * 24 + j
                w2d[:, k] = w[:, cm, kh, kw] / std[cm]
                mean_k[k], kh_of[k], kw_of[k] = mean[cm], kh, kw
        packed, rowscale = ops.pack_weights_i8(w2d.float(), 1, 192, 3)
        packed = packed.view(-1, 256)[:192].contiguous()                    # [3 planes x 64 channel slots, kpad = 256]
        dig = packed.view(3, 64, 256)[:, :cout, :192].double()
        wq = (dig[0] * 16384 + dig[1] * 128 + dig[2]) * rowscale.double()[:, None]      # the weights the kernel really uses
        tab = torch.zeros(4, 4, 4, 4, cout, dtype=torch.float64)
        for top in range(4):
            for bot in range(4):
                for lef in range(4):
                    for rig in range(4):
                        inb = (kh_of >= top) & (kh_of < 7 - bot) & (kw_of >= lef) & (kw_of < 7 - rig)
                        m = (wq[:, inb] * mean_k[inb][None, :]).sum(1)
                        tab[top, bot, lef, rig] = t - s * m
        self.cout, self.chw = cout, chw
        self.packed = packed.to(device)
        self.scale = fold.f32(s * rowscale.double(), device)
        self.tab = fold.f32(tab.reshape(256, cout), device)

    def __call__(self, img_u8):
        return ops.stem_u8(img_u8, self.packed, self.scale, self.tab, Cout=self.cout)


def pad16(c):
    """Channel counts are padded to 16 so that every int8 activation row is a legal TMA row (16-byte strides)."""
    return (c + 15) // 16 * 16


def _pad_channels(w2d, s, t, taps, cin_pad=None, cout_pad=None):
    """Zero-pad input channels (per tap) and output channels; padded outputs compute exactly 0 -> level 0."""
    cout = w2d.shape[0]
    cin = w2d.shape[1] // taps
    cin_pad, cout_pad = cin_pad or cin, cout_pad or cout
    if cin_pad == cin and cout_pad == cout:
        return w2d, s, t, cin
    w = torch.zeros(cout_pad, taps, cin_pad, dtype=w2d.dtype)
    w[:cout, :, :cin] = w2d.reshape(cout, taps, cin)
    s2, t2 = torch.zeros(cout_pad, dtype=s.dtype), torch.zeros(cout_pad, dtype=t.dtype)
    s2[:cout], t2[:cout] = s, t
    return w.reshape(cout_pad, taps * cin_pad), s2, t2, cin_pad


def _conv_gemm(sd, conv_key, bn_key, device, k=1, stride=1, pad=0, extra_scale=None, cin_pad=None, cout_pad=None):
    w = sd[conv_key + ".weight"]
    s, t = fold.conv_bn(sd, conv_key, bn_key, extra_scale)
    w2d, s, t, cin = _pad_channels(fold.w_khwc(w), s, t, k * k, cin_pad, cout_pad)
    gm = Gemm(w2d, s, t, cin, k, stride, pad, device)
    gm.alg_macs = int(w.shape[0]) * int(w[0].numel())          # the reference's (unpadded) Cout * Cin * k * k
    return gm


def head_pad(d):
    """Per-head channel count kept in HBM: a multiple of 4 so that every head slice is 32-bit aligned for the
    attention kernels (stage 4: 360 channels = 8 heads x 45 -> 8 x 48)."""
    return (d + 3) // 4 * 4


def _head_index(heads, d, dp):
    """positions of the reference's channels h*d + j inside the head-padded layout h*dp + j"""
    return (torch.arange(heads)[:, None] * dp + torch.arange(d)[None, :]).reshape(-1)


def _rep_gemm(sd, keys, device, cin_pad=None, cout_pad=None, out_heads=None, in_heads=None):
    """One dense 3x3 for one or several RepConv(+BN) branches sharing their input (outputs concatenated).
    out_heads = (heads, d, dp): every branch's output channels are scattered into the head-padded layout;
    in_heads: the same for the input channels (zero weights / zero outputs at the padding positions)."""
    ws, bs = zip(*(fold.repconv_dense3x3(sd, k) for k in keys))
    if out_heads is not None:
        heads, d, dp = out_heads
        idx = _head_index(heads, d, dp)
        ws2, bs2 = [], []
        for w, b in zip(ws, bs):
            w2 = torch.zeros(heads * dp, w.shape[1], dtype=w.dtype); w2[idx] = w
            b2 = torch.zeros(heads * dp, dtype=b.dtype); b2[idx] = b
            ws2.append(w2); bs2.append(b2)
        ws, bs = ws2, bs2
    w, b = torch.cat(ws, 0), torch.cat(bs, 0)
    if in_heads is not None:
        heads, d, dp = in_heads
        idx = _head_index(heads, d, dp)
        w3 = w.reshape(w.shape[0], 9, heads * d)
        w4 = torch.zeros(w.shape[0], 9, heads * dp, dtype=w.dtype)
        w4[:, :, idx] = w3
        w = w4.reshape(w.shape[0], -1)
    w, s, b, cin = _pad_channels(w, torch.ones_like(b), b, 9, cin_pad, cout_pad)
    gm = Gemm(w, s, b, cin, 3, 1, 1, device)
    # RepConv (sdtv2.py:111-132): 1x1 (C*C) + depthwise 3x3 (9*C) + 1x1 (C*C) MACs per pixel and branch
    gm.alg_macs = sum(2 * int(sd[k + ".0.body.0.weight"].shape[0]) * int(sd[k + ".0.body.0.weight"].shape[1]) +
                      9 * int(sd[k + ".0.body.2.0.weight"].shape[0]) for k in keys)
    return gm


def _dw(sd, conv_key, bn_key, device):
    w = sd[conv_key + ".weight"]
    if bn_key is None:
        return Dw(w, None, None, device)
    s, t = fold.bn_affine(sd, bn_key)
    return Dw(w, s, t, device)


class BackbonePlan:
    def __init__(self, model, device):
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        L = self.layers = {}
        dev = device
        L["stem"] = _conv_gemm(sd, "downsample1_1.encode_conv", "downsample1_1.encode_bn", dev, 7, 2, 3)
        for name in ("ConvBlock1_1.0", "ConvBlock1_2.0", "ConvBlock2_1.0", "ConvBlock2_2.0"):
            L[name + ".pw1"] = _conv_gemm(sd, name + ".Conv.pwconv1", name + ".Conv.bn1", dev)
            L[name + ".dw"] = _dw(sd, name + ".Conv.dwconv", None, dev)
            L[name + ".pw2"] = _conv_gemm(sd, name + ".Conv.pwconv2", name + ".Conv.bn2", dev)
            if SepDwPw.eligible(sd, name + ".Conv.dwconv", name + ".Conv.pwconv2"):
                L[name + ".dwpw"] = SepDwPw(sd, name + ".Conv.dwconv", name + ".Conv.pwconv2", name + ".Conv.bn2", dev)
            L[name + ".conv1"] = _conv_gemm(sd, name + ".conv1", name + ".bn1", dev, 3, 1, 1)
            L[name + ".conv2"] = _conv_gemm(sd, name + ".conv2", name + ".bn2", dev, 3, 1, 1)
        # every stage width is kept at a multiple of 16 channels in HBM (stage 4: 360 -> 368, zero padded)
        e = model.embed_dim
        self.width = {"block3": e[2], "block4": e[3]}
        for name, st in (("downsample1_2", 2), ("downsample2", 2), ("downsample3", 2), ("downsample4", 1)):
            L[name] = _conv_gemm(sd, name + ".encode_conv", name + ".encode_bn", dev, 3, st, 1,
                                 cout_pad=pad16(e[3]) if name == "downsample4" else None)
        nh = model.num_heads
        for name in [f"block3.{j}" for j in range(6)] + [f"block4.{j}" for j in range(2)]:
            c = self.width[name.split(".")[0]]
            cp = pad16(c)
            d = c // nh
            hp = (nh, d, head_pad(d)) if head_pad(d) != d else None      # q|k|v and the attention output are head-padded
            L[name + ".qkv"] = _rep_gemm(sd, [name + ".attn.q_conv", name + ".attn.k_conv", name + ".attn.v_conv"], dev,
                                         cin_pad=cp, out_heads=hp)
            L[name + ".proj"] = _rep_gemm(sd, [name + ".attn.proj_conv"], dev, cin_pad=None if hp else cp, cout_pad=cp,
                                          in_heads=hp)
            L[name + ".fc1"] = _conv_gemm(sd, name + ".mlp.fc1_conv", name + ".mlp.fc1_bn", dev, cin_pad=cp)
            L[name + ".fc2"] = _conv_gemm(sd, name + ".mlp.fc2_conv", name + ".mlp.fc2_bn", dev, cout_pad=cp)


class PixelDecoderPlan:
    def __init__(self, model, device):
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        L = self.layers = {}
        dev = device
        self.num_layers = model.encoder_cfg["num_layers"]
        sa = model.encoder_cfg["layer_cfg"]["self_attn_cfg"]
        self.group, self.dw_k = sa["group"], sa["dw_kernel_size"]
        self.cin_last = pad16(model.in_channels[-1])
        L["in_proj"] = _conv_gemm(sd, "encoder_in_proj.0", "encoder_in_proj.1", dev, cin_pad=self.cin_last)
        L["out_proj"] = _conv_gemm(sd, "encoder_out_proj.0", "encoder_out_proj.1", dev)

        def sepconv(prefix, key, gamma):
            L[prefix + ".pw1"] = _conv_gemm(sd, key + ".pwconv1.0", key + ".pwconv1.1", dev)
            L[prefix + ".dw"] = _dw(sd, key + ".dwconv.0", key + ".dwconv.1", dev)
            L[prefix + ".pw2"] = _conv_gemm(sd, key + ".pwconv2.0", key + ".pwconv2.1", dev, extra_scale=gamma)

        self.gamma3 = []
        for l in range(self.num_layers):
            k = f"encoder.layers.{l}"
            sepconv(f"{l}.conv", k + ".Conv", sd[k + ".gamma1"])
            sepconv(f"{l}.dcn_in", k + ".dcn.input_proj", None)
            sepconv(f"{l}.dcn_out", k + ".dcn.output_proj", sd[k + ".gamma2"])
            L[f"{l}.dcn_dw"] = _dw(sd, k + ".dcn.dw_conv.0", k + ".dcn.dw_conv.1", dev)
            L[f"{l}.offset"] = _conv_gemm(sd, k + ".dcn.offset.0", k + ".dcn.offset.1", dev)
            L[f"{l}.mask"] = _conv_gemm(sd, k + ".dcn.mask.0", k + ".dcn.mask.1", dev)
            L[f"{l}.fc1"] = _conv_gemm(sd, k + ".ffn.fc1_conv", k + ".ffn.fc1_bn", dev)
            L[f"{l}.fc2"] = _conv_gemm(sd, k + ".ffn.fc2_conv", k + ".ffn.fc2_bn", dev)
            self.gamma3.append(fold.f32(sd[k + ".gamma3"], dev))
        for i in range(model.num_inputs - 1):
            L[f"lateral.{i}"] = _conv_gemm(sd, f"lateral_convs.{i}.0", f"lateral_convs.{i}.1", dev)
            L[f"output.{i}"] = _dw(sd, f"output_convs.{i}.0", f"output_convs.{i}.1", dev)
        self.fpn_f16 = {i: FpnMergeF16(L[f"lateral.{i}"], dev) for i in range(model.num_inputs - 1)
                        if FpnMergeF16.eligible(L[f"lateral.{i}"])}
        L["mask_feature"] = _conv_gemm(sd, "mask_feature", None, dev)
        # einsum(mask_embed, mask_feature(s)) == (mask_embed W_mf) s + mask_embed b: the per-image product
        # [nq, C] x [W_mf^T ; b] is a tiny GEMM, after which mask_feature never has to be materialised.
        wmf = sd["mask_feature.weight"][:, :, 0, 0].double()                      # [C_out, C_in]
        waug = torch.cat([wmf.t(), sd["mask_feature.bias"].double()[None, :]], 0)  # [C_in + 1, C_out]
        L["mask_fold"] = Gemm(waug, torch.ones(waug.shape[0], dtype=torch.float64),
                              torch.zeros(waug.shape[0], dtype=torch.float64), waug.shape[1], device=dev)


class HeadPlan:
    def __init__(self, model, device):
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items() if not k.startswith("pixel_decoder.")}
        L = self.layers = {}
        dev = device
        cfg = model.transformer_decoder_cfg
        self.num_layers = cfg["num_layers"]
        self.heads = cfg["layer_cfg"]["self_attn_cfg"]["num_heads"]
        self.dim = cfg["layer_cfg"]["self_attn_cfg"]["embed_dims"]
        for l in range(self.num_layers):
            k = f"transformer_decoder.layers.{l}"
            for blk in ("cross_attn", "self_attn"):
                for p in ("q", "k", "v", "out"):
                    L[f"{l}.{blk}.{p}"] = _conv_gemm(sd, f"{k}.{blk}.attn.{p}_conv.0", f"{k}.{blk}.attn.{p}_conv.1", dev)
            L[f"{l}.fc1"] = _conv_gemm(sd, k + ".ffn.fc1", k + ".ffn.bn1", dev)
            L[f"{l}.fc2"] = _conv_gemm(sd, k + ".ffn.fc2", k + ".ffn.bn2", dev)
        self.query_feat = fold.f32(sd["query_feat.weight"], dev)
        self.query_embed = fold.f32(sd["query_embed.weight"], dev)
        self.level_embed = fold.f32(sd["level_embed.weight"], dev)
        L["cls"] = _conv_gemm(sd, "cls_embed", None, dev)
        L["me.fc1"] = _conv_gemm(sd, "mask_embed.fc1", None, dev)
        L["me.fc2"] = _conv_gemm(sd, "mask_embed.fc2", None, dev)
        L["me.out"] = _conv_gemm(sd, "mask_embed.fc_out", None, dev)
        # shortcut: Conv1d(nq -> nq) over the query axis + BN1d(nq), times the learnable scalar w
        wsc = float(sd["w"].item())
        s, t = fold.bn_affine(sd, "shortcut_conv.1")
        nq = sd["shortcut_conv.0.weight"].shape[0]
        L["shortcut"] = Gemm(sd["shortcut_conv.0.weight"][:, :, 0], s * wsc, t * wsc, nq, device=dev)
        self.pe_cache = {}
        self.num_feats = model.positional_encoding_cfg["num_feats"]

    def sine_pe(self, h, w, device):
        """SinePositionalEncoding (normalize=True, all-valid mask): positional_encoding.py:60-104 -> [h*w, 2F]."""
        key = (h, w)
        if key not in self.pe_cache:
            F_, temp, scale, eps = self.num_feats, 10000, 2 * math.pi, 1e-6
            ones = torch.ones(1, h, w, dtype=torch.int)
            y_e = ones.cumsum(1, dtype=torch.float32)
            x_e = ones.cumsum(2, dtype=torch.float32)
            y_e = (y_e + 0.0) / (y_e[:, -1:, :] + eps) * scale
            x_e = (x_e + 0.0) / (x_e[:, :, -1:] + eps) * scale
            dim_t = torch.arange(F_, dtype=torch.float32)
            dim_t = temp ** (2 * (dim_t // 2) / F_)
            px, py = x_e[:, :, :, None] / dim_t, y_e[:, :, :, None] / dim_t
            px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).view(1, h, w, -1)
            py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).view(1, h, w, -1)
            self.pe_cache[key] = torch.cat((py, px), dim=3).reshape(h * w, 2 * F_).contiguous().to(device)
        return self.pe_cache[key]


def plan_of(model, kind):
    """The model's device-side plan, rebuilt when its parameters changed since the plan was made (models._Engined)."""
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("spike2former_b200: move the model to a CUDA device first (no CPU path)")
    return model.current_plan(kind, dev)


# ------------------------------------------------------------------------------------------------ probing
class NullProbe:
    """Production probe: does nothing.  Layout tags tell a test probe how the reference lays the tensor out:
    'cm'  reference is channel-major [n, C, *spatial], ours is channels-last [n, *spatial, C];
    'same' identical flat order;  'reint_T' our buffer holds the reference's [C, nq] reinterpretation transposed;
    ('cm_heads', heads, d, dp): channels-last with every head's d channels padded to dp (stage 4: 45 -> 48).

    `active` probes (teacher forcing, census) may replace tensors and make the engine materialise what its fusions
    hide; an `observe` probe (oracle/probe.py::ObserverProbe) only LOOKS: the engine takes exactly the production
    code path (fused FPN merge, strided q|k|v, side stream, last-only SDME, CUDA-graph capture) and merely hands the
    probe every tensor that path produces anyway."""
    active = False
    observe = False

    def spike(self, name, t, layout="cm"):
        return t

    def real(self, name, t, layout="cm"):
        return t

    def scoped(self, prefix):
        return self


NOPROBE = NullProbe()


class FiringCensus:
    """Firing-rate / energy census (tools/cal_firing_num.py:140-171 of the reference hooks every neuron and averages
    its output): a probe that histograms the int8 levels of every neuron with s2f_level_hist -- one pass over 1 byte
    per neuron, no fp32 spike tensor.  `report()` -> {neuron: elements, hist[0..8], firing_rate (non-zero fraction),
    mean_level (the reference's mean output x 8)}.  Being a probe it disables the fusions that hide a neuron's
    levels from the host (like the teacher-forcing probe of the tests), so it is an analysis mode, not the fast path."""
    active = True

    def __init__(self, prefix="", table=None, keep=False, kept=None):
        self.prefix, self.table, self.keep = prefix, ({} if table is None else table), keep
        self.kept = {} if kept is None else kept

    def scoped(self, prefix):
        return FiringCensus(self.prefix + prefix, self.table, self.keep, self.kept)

    def spike(self, name, t, layout="cm"):
        full = self.prefix + name
        tc = t if t.is_contiguous() else t.contiguous()
        self.table[full] = (ops.level_hist(tc, self.table[full][0] if full in self.table else None),
                            (self.table[full][1] if full in self.table else 0) + t.numel())
        if self.keep:
            self.kept[full] = tc.clone()
        return t

    def real(self, name, t, layout="cm"):
        return t

    def report(self):
        out = {}
        for name, (hist, n) in self.table.items():
            h = hist.cpu().tolist()
            fired = n - h[0]
            out[name] = dict(elements=n, hist=h, firing_rate=fired / max(n, 1),
                             mean_level=sum(i * c for i, c in enumerate(h)) / max(n, 1))
        return out


# ------------------------------------------------------------------------------------------------ backbone
def _conv_block(L, name, s, sp, n, H, W, pr):
    """MS_ConvBlock (sdtv2.py:207-219) on stream s with spike twin sp = LIF(s)."""
    sp = pr.spike(f"{name}.Conv.spike1", sp)
    _, a = L[name + ".pw1"](sp, n, H, W, spike=True, f32=False)
    a = pr.spike(f"{name}.Conv.spike2", a)
    if name + ".dwpw" in L and a.dtype == torch.int8:
        s2, sp2 = L[name + ".dwpw"](a, n, H, W, residual=s, f32=True, spike=True)
    else:
        d, _ = L[name + ".dw"](a, n, H, W, f32=True)
        s2, sp2 = L[name + ".pw2"](d, n, H, W, residual=s, f32=True, spike=True)
    s2 = pr.real(f"{name}.spike1", s2)
    sp2 = pr.spike(f"{name}.spike1", sp2)
    _, a = L[name + ".conv1"](sp2, n, H, W, spike=True)
    a = pr.spike(f"{name}.spike2", a)
    return L[name + ".conv2"](a, n, H, W, residual=s2, f32=True, spike=True)


def _ms_block(L, name, s, sp, n, H, W, heads, pr, C):
    """MS_Block (sdtv2.py:298-383): SDSA with the RepConv branches as one dense 3x3, then MS_MLP.
    C is the reference width; the stream / spike buffers carry pad16(C) channels (zeros beyond C)."""
    CP = s.shape[-1]
    d = C // heads
    dp = head_pad(d)
    CA = heads * dp                                                        # width of q, k, v and of the attention output
    N = H * W
    sp = pr.spike(f"{name}.attn.head_spike", sp)
    _, qkv = L[name + ".qkv"](sp, n, H, W, spike=True)                      # [n,H,W,3*CA] levels of q|k|v
    if pr.active:
        def ref_layout(t):                                                  # head-padded [.., heads*dp] -> reference [.., C]
            return t.reshape(n, H, W, heads, dp)[..., :d].reshape(n, H, W, C).contiguous()

        def padded(t):
            out = torch.zeros(n, H, W, heads, dp, dtype=t.dtype, device=t.device)
            out[..., :d] = t.reshape(n, H, W, heads, d)
            return out.reshape(n, H, W, CA)

        q, k, v = (padded(pr.spike(f"{name}.attn.{nm}_spike", ref_layout(qkv[..., i * CA:(i + 1) * CA])))
                   for i, nm in enumerate("qkv"))
        ld = CA
    else:
        q, k, v = qkv[..., :CA], qkv[..., CA:2 * CA], qkv[..., 2 * CA:]
        ld = 3 * CA
        if pr.observe:
            for t, nm in ((q, "q"), (k, "k"), (v, "v")):
                pr.spike(f"{name}.attn.{nm}_spike", t, ("cm_heads", heads, d, dp))
    out_w = CA if dp != d else CP
    att, _ = ops.linear_attn(q, k, v, n=n, Nq=N, Nk=N, heads=heads, d=dp, q_ld=ld, kv_ld=ld, out_ld=out_w,
                             out_scale=(d ** -0.5) * INV ** 3)
    att = att.view(n, H, W, out_w)
    if pr.active:
        if dp != d:
            att = padded(pr.spike(f"{name}.attn.attn_spike", ref_layout(att)))
        else:
            att = pr.spike(f"{name}.attn.attn_spike", att)
    elif pr.observe:
        pr.spike(f"{name}.attn.attn_spike", att, ("cm_heads", heads, d, dp) if dp != d else "cm")
    s2, sp2 = L[name + ".proj"](att, n, H, W, residual=s, f32=True, spike=True)
    s2 = pr.real(f"{name}.mlp.fc1_spike", s2)
    sp2 = pr.spike(f"{name}.mlp.fc1_spike", sp2)
    _, a = L[name + ".fc1"](sp2, n, H, W, spike=True)
    a = pr.spike(f"{name}.mlp.fc2_spike", a)
    return L[name + ".fc2"](a, n, H, W, residual=s2, f32=True, spike=True)


def backbone_forward(model, img, probe=NOPROBE, pre=None):
    """Spiking_vit_MetaFormer.forward_features (sdtv2.py:614-651).  img fp32 [B,3,H,W] (NCHW, as the reference
    receives it).  Returns [(stream fp32 [n,h,w,C], levels int8 [n,h,w,C])] for x1..x4."""
    plan = plan_of(model, BackbonePlan)
    L, pr = plan.layers, probe
    B, cin, H, W = img.shape
    n = B * model.T
    if img.dtype == torch.uint8:
        # uint8 batch: data preprocessor + stem in one tensor-core launch
        chw = img.shape[1] == 3 and img.shape[3] != 3
        if not chw:
            B, H, W, cin = img.shape
        key = ("stem_u8", chw, bool(pre.channel_conversion), tuple(pre._mean_host), tuple(pre._std_host))
        if key not in L:
            sd = {k: v.detach().cpu() for k, v in model.state_dict().items() if k.startswith("downsample1_1.")}
            L[key] = StemU8(sd, pre._mean_host, pre._std_host, pre.channel_conversion, chw, img.device)
        n = B * model.T
        s, sp = L[key](img)
    elif model.T == 1 and cin == 3 and img.is_contiguous() and img.dtype == torch.float32:
        # the planar NCHW batch as the caller holds it: the stem kernel reads it through pixel / channel strides, so the
        # NHWC copy of the image is never made
        s, sp = L["stem"](img, n, H, W, f32=True, spike=True, a_stride_m=1, a_stride_k=H * W, a_img_stride=cin * H * W)
    else:
        x = img.permute(0, 2, 3, 1).contiguous()       # a no-op for channels-last memory (SegDataPreProcessor output)
        if model.T > 1:
            x = x.unsqueeze(0).expand(model.T, -1, -1, -1, -1).reshape(n, H, W, cin).contiguous()
        s, sp = L["stem"](x, n, H, W, f32=True, spike=True)
    H, W = H // 2, W // 2
    feats = []

    def down(name, s, sp, H, W, stride):
        s = pr.real(f"{name}.encode_spike", s)
        sp = pr.spike(f"{name}.encode_spike", sp)
        return L[name](sp, n, H, W, f32=True, spike=True)

    s = pr.real("ConvBlock1_1.0.Conv.spike1", s)
    s, sp = _conv_block(L, "ConvBlock1_1.0", s, sp, n, H, W, pr)
    feats.append((s, sp))
    s, sp = down("downsample1_2", s, sp, H, W, 2); H, W = H // 2, W // 2
    s = pr.real("ConvBlock1_2.0.Conv.spike1", s)
    s, sp = _conv_block(L, "ConvBlock1_2.0", s, sp, n, H, W, pr)
    feats.append((s, sp))
    s, sp = down("downsample2", s, sp, H, W, 2); H, W = H // 2, W // 2
    s = pr.real("ConvBlock2_1.0.Conv.spike1", s)
    s, sp = _conv_block(L, "ConvBlock2_1.0", s, sp, n, H, W, pr)
    s = pr.real("ConvBlock2_2.0.Conv.spike1", s)
    s, sp = _conv_block(L, "ConvBlock2_2.0", s, sp, n, H, W, pr)
    feats.append((s, sp))
    s, sp = down("downsample3", s, sp, H, W, 2); H, W = H // 2, W // 2
    for j in range(6):
        s = pr.real(f"block3.{j}.attn.head_spike", s)
        s, sp = _ms_block(L, f"block3.{j}", s, sp, n, H, W, model.num_heads, pr, plan.width["block3"])
    s, sp = down("downsample4", s, sp, H, W, 1)
    for j in range(2):
        s = pr.real(f"block4.{j}.attn.head_spike", s)
        s, sp = _ms_block(L, f"block4.{j}", s, sp, n, H, W, model.num_heads, pr, plan.width["block4"])
    feats.append((s, sp))
    return feats


def export_backbone_feats(model, feats):
    """Reference output contract (sdtv2.py:639-651): 'Qsnn' -> [T,B,C,H,W]; 'snn' -> mean over T; else [T*B,C,H,W]."""
    T = model.T
    outs = []
    widths = [model.embed_dim[0] // 2, model.embed_dim[0], model.embed_dim[1], model.embed_dim[3]]
    for (s, sp), c in zip(feats, widths):
        n, h, w, _ = s.shape
        t = s[..., :c].permute(0, 3, 1, 2)
        if model.decode_mode == "Qsnn":
            t = t.reshape(T, n // T, c, h, w)
        elif model.decode_mode == "snn":
            t = t.reshape(T, n // T, c, h, w).mean(0, keepdim=True)
        t._s2f = (s, sp)
        outs.append(t)
    return outs


def _import_feats(x):
    """Public-API inputs ([T,B,C,H,W] tensors) -> internal (stream, levels) pairs."""
    feats = []
    for t in x:
        cached = getattr(t, "_s2f", None)
        if cached is not None:
            feats.append(cached)
            continue
        if t.dim() == 5:
            t = t.flatten(0, 1)
        s = t.permute(0, 2, 3, 1).contiguous().float()
        sp, _, _ = ops.nilif(s)
        feats.append((s, sp))
    return feats


# ------------------------------------------------------------------------------------------------ pixel decoder
def _sepconv_spike(L, prefix, key, sp, n, H, W, pr, residual=None, want_spike=True):
    """SepConv_Spike (SNN_core.py:47-63) from the spike twin of its input; pw2 epilogue adds the residual."""
    sp = pr.spike(key + ".spike1", sp)
    _, a = L[prefix + ".pw1"](sp, n, H, W, spike=True)
    a = pr.spike(key + ".spike2", a)
    _, a = L[prefix + ".dw"](a, n, H, W, spike=True)
    a = pr.spike(key + ".spike3", a)
    return L[prefix + ".pw2"](a, n, H, W, residual=residual, f32=True, spike=want_spike)


OVERLAP_FPN_TAIL = _os.environ.get("S2F_OVERLAP", "1") != "0"
FUSE_STEM_U8 = _os.environ.get("S2F_FUSE_STEM", "1") != "0"      # uint8 input: preprocessor folded into the tensor-core stem


def pixel_decoder_forward(model, feats, probe=NOPROBE, want_mask_feature=True, side_stream=None):
    """DCNTransformerEncoderPixelDecoder.forward (pixel_decoder.py:417-472).
    Returns (mask_feature fp32 [n,H1,W1,C], memory levels, [y32, y64, y128] fp32 channels-last).
    side_stream: the finest FPN level (lateral + merge + output conv at H/2 x W/2, which only feeds mask_feature) is
    enqueued there, so that it runs beside the transformer decoder's many small launches; the caller joins the stream
    before it touches the first return value."""
    plan = plan_of(model, PixelDecoderPlan)
    L, pr = plan.layers, probe
    pd = "pixel_decoder."
    s4, sp4 = feats[-1]
    n, H, W, _ = s4.shape
    C = model.feat_channels
    G = plan.group
    if sp4.shape[-1] != plan.cin_last:            # public-API input with the reference's 360 channels
        sp4 = torch.nn.functional.pad(sp4, (0, plan.cin_last - sp4.shape[-1]))
    sp4 = pr.spike(pd + "last_feat_conv_spike", sp4)
    q, qs = L["in_proj"](sp4, n, H, W, f32=True, spike=True)
    for l in range(plan.num_layers):
        k = pd + f"encoder.layers.{l}"
        q = pr.real(k + ".Conv.spike1", q)
        q1, q1s = _sepconv_spike(L, f"{l}.conv", k + ".Conv", qs, n, H, W, pr, residual=q)
        # ---- DCNv3 (dcnv3.py:198-233)
        q1 = pr.real(k + ".dcn.input_proj.spike1", q1)
        x, _ = _sepconv_spike(L, f"{l}.dcn_in", k + ".dcn.input_proj", q1s, n, H, W, pr, want_spike=False)
        x = pr.real(k + ".dcn.x", x, "same")
        a = pr.spike(k + ".dcn.dw_spike", q1s)
        _, a = L[f"{l}.dcn_dw"](a, n, H, W, spike=True)
        a = pr.spike(k + ".dcn.offset_spike", a)
        # NCHW conv outputs reinterpreted as [n,H,W,-1]: written channel-major, read back channels-last
        off, _ = L[f"{l}.offset"](a, n, H, W, f32=True, transposed=True)
        off = pr.real(k + ".dcn.offset", off.view(n, H, W, -1), "same")
        _, msk = L[f"{l}.mask"](a, n, H, W, spike=True, transposed=True)
        msk = pr.spike(k + ".dcn.mask_spike", msk.view(n, H, W, -1), "same")
        g = ops.dcnv3_gather(x, off, msk, n=n, H=H, W=W, G=G, Cg=C // G, K=3, offset_scale=1.0)
        g = pr.real(k + ".dcn.output_proj.spike1", g)
        gs, _, _ = ops.nilif(g)
        q2, q2s = _sepconv_spike(L, f"{l}.dcn_out", k + ".dcn.output_proj", gs, n, H, W, pr, residual=q1)
        # ---- MS_MLP with the reinterpreted output (mmcv_spike/transformer.py:817-831)
        q2 = pr.real(k + ".ffn.fc1_spike", q2)
        a = pr.spike(k + ".ffn.fc1_spike", q2s)
        _, a = L[f"{l}.fc1"](a, n, H, W, spike=True)
        a = pr.spike(k + ".ffn.fc2_spike", a)
        r, _ = L[f"{l}.fc2"](a, n, H, W, f32=True, transposed=True)
        q, qs = ops.affine_add_lif(r.view(n, H, W, C), plan.gamma3[l], q2)
    q = pr.real(pd + "encoder_out_proj_spike", q)
    memory = pr.spike(pd + "encoder_out_proj_spike", qs)
    y, _ = L["out_proj"](memory, n, H, W, f32=True)
    y = pr.real(pd + "y0", y)
    outs = [y]
    hp, wp = H, W
    forked = False
    for i in range(model.num_inputs - 2, -1, -1):
        forked = side_stream is not None and i == 0 and not pr.active and not want_mask_feature
        if forked:
            side_stream.wait_stream(torch.cuda.current_stream())
        with (torch.cuda.stream(side_stream) if forked else contextlib.nullcontext()):
            s_i, sp_i = feats[i]
            _, h, w, _ = s_i.shape
            sp_i = pr.spike(pd + f"lateral_convs_spike.{i}", sp_i)
            lat = L[f"lateral.{i}"]
            if not pr.active and lat.tc is not None and n * h * w >= TC_MIN_ROWS and C % 16 == 0:
                # lateral 1x1 + BN, bilinear upsample of the coarser level, add and NI-LIF in one launch: the fp32 lateral
                # map (1 GB at batch 16 for the 256^2 level) is never written
                fm = plan.fpn_f16.get(i)
                if fm is not None and h == 2 * hp and w == 2 * wp and sp_i.dtype == torch.int8:
                    _, ys = fm(sp_i, y, n, h, w)
                else:
                    _, ys = lat(sp_i, n, h, w, spike=True, up_prev=y)
            else:
                cur, _ = lat(sp_i, n, h, w, f32=True)
                ys, yf = ops.upsample_add_lif(cur, y, n=n, H=h, W=w, Hp=hp, Wp=wp, C_=C, want_f32=pr.active)
                if pr.active:
                    pr.real(pd + f"output_convs_spike.{i}", yf)
            ys = pr.spike(pd + f"output_convs_spike.{i}", ys)
            # the finest level only feeds mask_feature_spike: its fp32 map (1 GB at batch 16) is not materialised
            y, ysp = L[f"output.{i}"](ys, n, h, w, f32=(i > 0 or pr.active), spike=(i == 0))
            if y is not None:
                y = pr.real(pd + f"y{model.num_inputs - 1 - i}", y)
            outs.append(y)
            hp, wp = h, w
    if pr.active:
        pr.real(pd + "mask_feature_spike", y)
    with (torch.cuda.stream(side_stream) if forked else contextlib.nullcontext()):   # an observer's tap copies on the producing stream
        ysp = pr.spike(pd + "mask_feature_spike", ysp)
    mf = None
    if want_mask_feature or pr.active:
        mf, _ = L["mask_feature"](ysp, n, hp, wp, f32=True)
        mf = pr.real(pd + "mask_feature", mf)
    return (mf if want_mask_feature else ysp), memory, outs[:3]


def pixel_decoder_forward_public(model, x):
    mf, memory, outs = pixel_decoder_forward(model, _import_feats(x))
    T = x[0].shape[0] if x[0].dim() == 5 else 1

    def nchw(t):
        n, h, w, c = t.shape
        return t.permute(0, 3, 1, 2).reshape(T, n // T, c, h, w)

    return nchw(mf), nchw(memory.float() * INV), [nchw(o) for o in outs]


# ------------------------------------------------------------------------------------------------ decoder + SDME
def _attention(L, key, pfx, q_sp, k_sp, v_sp, n, nq, nk, heads, dim, residual, pr, kv_cache=None, attn_mask=None):
    """{Cross,}MultiHeadAttentionBlock (mmcv_spike/transformer.py:237-278, 318-361) from the LIF levels of
    its three inputs.  Linear form: NI-LIF(Q (K^T V) / sqrt(dim)); out_conv epilogue adds the residual.
    attn_mask (bool [n*heads, nq, nk], True = masked_fill(0), :265-269 / :349-353) switches to the key-walking kernel
    s2f_dec_attn, because a mask does not commute with the Q (K^T V) re-association."""
    q_sp = pr.spike(key + ".q_conv_spike", q_sp, "same")
    _, q = L[pfx + ".q"](q_sp, n, nq, 1, spike=True)
    q = pr.spike(key + ".q_spike", q.view(n, nq, dim), "same")
    k_sp = pr.spike(key + ".k_conv_spike", k_sp, "same")
    _, k = L[pfx + ".k"](k_sp, n, nk, 1, spike=True)
    k = pr.spike(key + ".k_spike", k.view(n, nk, dim), "same")
    v_sp = pr.spike(key + ".v_conv_spike", v_sp, "same")
    _, v = L[pfx + ".v"](v_sp, n, nk, 1, spike=True)
    v = pr.spike(key + ".v_spike", v.view(n, nk, dim), "same")
    if attn_mask is None:
        att, _ = ops.linear_attn(q, k, v, n=n, Nq=nq, Nk=nk, heads=heads, d=dim // heads,
                                 out_scale=INV ** 3 / (dim ** 0.5))
    else:
        att, _ = ops.dec_attn(q, k, v, n=n, Nq=nq, Nk=nk, heads=heads, d=dim // heads, mask=attn_mask,
                              out_scale=INV ** 3 / (dim ** 0.5))
    att = pr.spike(key + ".attn_spike", att, "same")
    out, _ = L[pfx + ".out"](att, n, nq, 1, residual=residual, f32=True)
    return out.view(n, nq, dim)


def head_forward(model, feats, probe=NOPROBE, last_only=False, cross_attn_masks=None):
    """mmdet MaskFormerHead.forward (dense_heads/maskformer_head.py:498-586).
    Returns (cls [L,n,nq,K+1], mask levels int8 [L,n,nq,C], mask_feature [n,h,w,C]); L = 1 when last_only.
    cross_attn_masks: optional list (one entry per decoder layer, None or bool [n*heads, nq, nk_level]) -- the
    `cross_attn_mask` argument of DetrTransformerDecoderLayer.forward (detr_layers.py:491-559); the Spike2Former configs
    pass None (maskformer_head.py:563)."""
    pd_model = model.pixel_decoder
    pr = probe
    side = None
    if OVERLAP_FPN_TAIL and not pr.active and ops._PROF is None:
        side = getattr(model, "_side_stream", None)
        if side is None or side.device != feats[0][0].device:
            side = model._side_stream = torch.cuda.Stream(device=feats[0][0].device)
    mf, memory, ms = pixel_decoder_forward(pd_model, feats, pr, want_mask_feature=False, side_stream=side)   # mf: mask_feature_spike levels
    plan = plan_of(model, HeadPlan)
    L = plan.layers
    n = mf.shape[0]
    nq, dim, heads = model.num_queries, plan.dim, plan.heads
    dev = mf.device
    hd = ""
    qe = plan.query_embed                                        # [nq, C]
    qf = plan.query_feat.unsqueeze(0).expand(n, nq, dim).contiguous()
    # key / value spikes per level: LIF(y + level_embed + pos) and LIF(y + level_embed) -- shared by layers i, i+3
    lvl_k, lvl_v, lvl_n = [], [], []
    for i in range(3):
        y = ms[i]
        _, h, w, _ = y.shape
        pe = plan.sine_pe(h, w, dev)
        ones = torch.ones(dim, dtype=torch.float32, device=dev)
        if y.numel() % 16 == 0 and dim % 4 == 0 and pe.numel() % 4 == 0:
            ksp, vsp = ops.nilif_pair(y, ones, plan.level_embed[i].contiguous(), pe, residual_period=pe.numel())
        else:
            ksp, _, _ = ops.nilif(y, scale=ones, shift=plan.level_embed[i].contiguous(), residual=pe,
                                  residual_period=pe.numel())
            vsp, _, _ = ops.nilif(y, scale=ones, shift=plan.level_embed[i].contiguous())
        lvl_k.append(ksp.view(n, h * w, dim)); lvl_v.append(vsp.view(n, h * w, dim)); lvl_n.append(h * w)
    states = [qf]
    for i in range(plan.num_layers):
        lv = i % 3
        k = f"transformer_decoder.layers.{i}"
        q_sp, _, _ = ops.nilif(qf, residual=qe, residual_period=qe.numel())
        qf = _attention(L, k + ".cross_attn.attn", f"{i}.cross_attn", q_sp, lvl_k[lv], lvl_v[lv], n, nq, lvl_n[lv], heads,
                        dim, qf, pr, attn_mask=cross_attn_masks[i] if cross_attn_masks is not None else None)
        qf = pr.real(k + ".self_attn.attn.v_conv_spike", qf, "same")
        qk_sp, _, _ = ops.nilif(qf, residual=qe, residual_period=qe.numel())
        v_sp, _, _ = ops.nilif(qf)
        qf = _attention(L, k + ".self_attn.attn", f"{i}.self_attn", qk_sp, qk_sp, v_sp, n, nq, nq, heads, dim, qf, pr)
        # ---- MSDA_FFN with both reinterpreting reshapes (mmcv_spike/transformer.py:768-784)
        qf = pr.real(k + ".ffn.fc1_spike", qf, "same")
        a, _, _ = ops.nilif(qf, transpose=(nq, dim))             # [C,nq] matrix stored as [nq(pos), C]
        a = pr.spike(k + ".ffn.fc1_spike", a, "reint_T")
        _, a = L[f"{i}.fc1"](a, n, nq, 1, spike=True)
        a = pr.spike(k + ".ffn.fc2_spike", a.view(n, nq, -1), "cm")
        r, _ = L[f"{i}.fc2"](a, n, nq, 1, f32=True, transposed=True)
        qf, _ = ops.affine_add_lif(r.view(n, nq, dim), None, qf, want_spike=False)
        qf = pr.real(k + ".out", qf, "same")
        states.append(qf)
    # ---- SDME (maskformer_head.py:571-582)
    if last_only and not pr.active:
        states = states[-1:]
    Ls = len(states)
    od = torch.stack(states)                                     # [L, n, nq, C]
    lv = ops.sigmoid_lif(od)
    lv = pr.spike("decoder_out_spike", lv, "same")
    pr.spike("shortcut_conv_spike", lv, "same")
    rows = Ls * n
    half = model.alpha * INV                                     # alpha * level / 8
    cls, _ = L["cls"](lv, rows, nq, 1, f32=True, a_scale=half)
    _, a = L["me.fc1"](lv, rows, nq, 1, spike=True, a_scale=half)
    a = pr.spike("mask_embed.spike1", a.view(Ls, n, nq, dim), "same")
    _, a = L["me.fc2"](a, rows, nq, 1, spike=True, a_scale=half)
    a = pr.spike("mask_embed.spike2", a.view(Ls, n, nq, dim), "same")
    # shortcut over the query axis: A[m=c][k=q] = lv[q*C + c]; output [q', c] = transposed store
    sc, _ = L["shortcut"](lv, rows, dim, 1, f32=True, transposed=True, a_scale=half, a_img_stride=nq * dim,
                          a_stride_m=1, a_stride_k=dim)
    m_f, me = L["me.out"](a, rows, nq, 1, residual=sc.view(rows, nq, 1, dim), f32=pr.active, spike=True, a_scale=half)
    if pr.active:
        pr.real("mask_embed_spike", m_f.view(Ls, n, nq, dim), "same")
    me = pr.spike("mask_embed_spike", me.view(Ls, n, nq, dim), "same")
    cls = cls.view(Ls, n, nq, -1)
    if side is not None:
        torch.cuda.current_stream().wait_stream(side)              # mask_feature_spike is complete from here on
        mf.record_stream(torch.cuda.current_stream())
    return cls, me, mf


def _mask_pred(model, me_rows, ysp, transposed):
    """einsum('bqc,bchw->bqhw', mask_embed, mask_feature) (maskformer_head.py:581) with the 1x1 mask_feature
    convolution folded into the per-image weights: mask = (m W_mf) s + m b, m = alpha*level/8.
    me_rows: int8 levels [n, R, C]; ysp: int8 levels of mask_feature_spike [n, h, w, C_in]."""
    pdl = plan_of(model.pixel_decoder, PixelDecoderPlan).layers
    n, R, dim = me_rows.shape
    _, h, w, cin = ysp.shape
    wb, _ = pdl["mask_fold"](me_rows.contiguous(), 1, n * R, 1, f32=True, a_scale=model.alpha * INV)   # [1, n*R, 1, cin+1]
    packed, sc, sh = ops.pack_rows_i8_device(wb.view(n * R, cin + 1), n_img=n, rows_per_img=R, K=cin, pieces=TC_PIECES,
                                             post_scale=INV, bias_col=cin)
    # the reference computes mask_feature (C_in x C_out per pixel, pixel_decoder.py:467-470) and then the einsum
    # (R x C_out per pixel, maskformer_head.py:581); folded, only R x C_in per pixel is executed
    out, _ = ops.gemm_tc(ysp, packed, n=n, H=h, W=w, Cin=cin, Cout=R, scale=sc, shift=sh, pieces=TC_PIECES,
                         want_f32=True, transposed=transposed, per_image=True, alg_macs=cin * dim + R * dim)
    return out


def head_forward_public(model, x):
    feats = _import_feats(x)
    T = x[0].shape[0] if x[0].dim() == 5 else 1
    cls, me, ysp = head_forward(model, feats)
    Ls, n, nq, dim = me.shape
    _, h, w, _ = ysp.shape
    rows = me.permute(1, 0, 2, 3).reshape(n, Ls * nq, dim)
    masks = _mask_pred(model, rows, ysp, True).view(n, Ls, nq, h, w).permute(1, 0, 2, 3, 4)
    B = n // T
    return cls.view(Ls, T, B, nq, -1).mean(1), masks.view(Ls, T, B, nq, h, w).mean(1)


def _predict_from(model, feats, img_shape, probe=NOPROBE, labels=False, T=1):
    cls, me, ysp = head_forward(model, feats, probe, last_only=True)
    n, h, w, _ = ysp.shape
    nq = model.num_queries
    mp = _mask_pred(model, me[-1], ysp, False)                              # [n, h, w, nq] pixel-major
    cl = cls[-1].contiguous()
    if T > 1:
        # `.mean(1)` over the time axis of the class scores and of the mask einsum (dense_heads/maskformer_head.py:
        # 574-582); only the T = 4 cocostuff configs reach this -- two small torch reductions, not a hot loop
        n //= T
        mp = mp.view(T, n, h, w, nq).mean(0)
        cl = cl.view(T, n, nq, -1).mean(0)
    mp = probe.real("mask_pred", mp)
    cl = probe.real("cls_score", cl, "same")
    logits, lab = ops.semantic_tail(mp.view(n, h * w, nq), cl, n=n, Q=nq, K=model.num_classes, h=h, w=w,
                                    H=img_shape[0], W=img_shape[1], want_logits=not labels, want_labels=labels)
    return lab if labels else logits


def head_predict(model, x, img_shape, probe=NOPROBE):
    """mmseg MaskFormerHead.predict (decode_heads/maskformer_head.py:138-180) -> [B, K, H, W]."""
    feats = _import_feats(x)
    T = x[0].shape[0] if x[0].dim() == 5 else 1
    return _predict_from(model, feats, img_shape, probe, T=T)


class GraphedForward:
    """One captured CUDA graph of segmentor_logits for a fixed input shape (see EncoderDecoder._run)."""

    def __init__(self, seg, example, labels, probe=NOPROBE):
        """probe: an observing probe (tests) whose taps become extra copy nodes of the captured graph."""
        if probe.active:
            raise RuntimeError("an active probe changes the code path and cannot be captured; use an observer")
        self.static_in = torch.empty_like(example)
        self.static_in.copy_(example)
        side = torch.cuda.Stream(device=example.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):                     # builds the plans, sets kernel attributes, fills the PE caches
                segmentor_logits(seg, self.static_in, labels=labels)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(example.device)
        self.graph = torch.cuda.CUDAGraph()
        l0 = ops.launch_count()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = segmentor_logits(seg, self.static_in, probe, labels=labels)
        # the graph replays raw device pointers into these plans: keep them alive as long as the graph
        self.plans = (seg.backbone._plan, seg.decode_head._plan, seg.decode_head.pixel_decoder._plan)
        self.launches = ops.launch_count() - l0    # kernels of this library inside one replay

    def __call__(self, x):
        if x.data_ptr() != self.static_in.data_ptr():
            self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out


HBM_CLASSES = ("nilif", "dwconv", "dcn_gather", "upsample_add_lif", "elementwise", "linear_attn")


def profile_dominant(seg, img, steps=2, labels=False):
    """Time every launch of `steps` forwards with CUDA events on the launch stream and report the kernel
    class that takes the largest share of the step (bench.py's `roofline` object) plus a per-class table:
    ms, launches, ALGORITHMIC TFLOP/s (the reference layers' work), executed TFLOP/s, GB/s."""
    prof = ops.Profiler()
    with torch.no_grad():
        segmentor_logits(seg, img, labels=labels)       # warm
        torch.cuda.synchronize()
        for _ in range(steps):
            # park the GPU behind a ~0.3 s spin kernel while the host enqueues the whole forward: the events then
            # bracket device time only, not the host's per-launch overhead (the production path replays a CUDA graph)
            torch.cuda._sleep(int(6e8))
            ops.set_profiler(prof)
            try:
                segmentor_logits(seg, img, labels=labels)
            finally:
                ops.set_profiler(None)
            torch.cuda.synchronize()
    agg = prof.summary()
    total = sum(a["ms"] for a in agg.values())
    name, top = max(agg.items(), key=lambda kv: kv[1]["ms"])
    tensor = name.startswith("gemm") or name == "semantic_tail"
    sec = top["ms"] / 1e3
    classes = {}
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        t = v["ms"] / 1e3
        classes[k] = dict(ms=round(v["ms"] / steps, 3), launches=v["launches"] // steps,
                          tflops_algorithmic=round(v["flops"] / t / 1e12, 1), tflops_executed=round(v["exec_flops"] / t / 1e12, 1),
                          gbs=round(v["bytes"] / t / 1e9, 0))
    return dict(kernel=name, bound="tensor" if tensor else "hbm",
                achieved=(top["flops"] / sec / 1e12) if tensor else (top["bytes"] / sec / 1e9),
                achieved_executed=(top["exec_flops"] / sec / 1e12) if tensor else None,
                unit="TFLOP/s" if tensor else "GB/s", share=top["ms"] / total, launches=top["launches"] // steps,
                dtype="f32 (int8 spikes x fp32 weights)" if name == "gemm_simt" else "int8 spikes x int8 weight digits",
                per_class_ms={k: v["ms"] for k, v in classes.items()}, classes=classes, step_ms=total / steps)


def segmentor_logits(seg, img, probe=NOPROBE, labels=False):
    """EncoderDecoder.encode_decode (encoder_decoder.py:125-133): internal tensors go straight to the head."""
    pre = None
    img_hw = tuple(img.shape[-2:])
    if img.dtype == torch.uint8:
        pre = getattr(seg, "data_preprocessor", None)
        if pre is None:
            raise RuntimeError("uint8 input needs a data_preprocessor (mean / std / bgr_to_rgb) on the segmentor")
        tc = pre.test_cfg or {}
        chw = img.shape[1] == 3 and img.shape[3] != 3
        h_, w_ = (img.shape[2], img.shape[3]) if chw else (img.shape[1], img.shape[2])
        cout0 = seg.backbone.embed_dim[0] // 2
        no_pad = pre._padded_size(int(h_), int(w_), tc.get("size"), tc.get("size_divisor")) == (int(h_), int(w_))
        fused = (FUSE_STEM_U8 and pre._enable_normalize and no_pad and
                 seg.backbone.T == 1 and min(h_, w_) >= 8 and cout0 % 16 == 0 and cout0 <= 64 and img.is_contiguous())
        if not fused:
            # separate preprocessing kernel (padding requested, or a stem width the fused kernel does not cover)
            img = pre.normalized(img, tc.get("size", None), tc.get("size_divisor", None)).permute(0, 3, 1, 2)
            pre = None
            img_hw = tuple(img.shape[-2:])
        else:
            img_hw = (int(h_), int(w_))
    scoped = probe.active or probe.observe
    feats = backbone_forward(seg.backbone, img, probe.scoped("backbone.") if scoped else probe, pre=pre)
    pr = probe.scoped("decode_head.") if scoped else probe
    return _predict_from(seg.decode_head, feats, img_hw, pr, labels, T=seg.backbone.T)
