"""Parameter trees with the reference's state_dict names.

The reference's checkpoint keys are the compatibility surface:
a reference `state_dict` must load with `strict=True`.  The modules here are
*holders* only -- plain torch layers hung at the reference's dotted paths --
and are never called; the forward pass is executed by engine.py on the CUDA
kernels.  Each builder cites the reference constructor it mirrors.
"""
from __future__ import annotations

import torch.nn as nn


class NeuronSlot(nn.Module):
    """Placeholder at the dotted path of a reference `Q_IFNode`.

    Keeps the neuron addressable by name (tools/cal_firing_num.py:140-171 walks
    named_modules) and honours the duck-typed reset protocol of
    resetmodel_hook.py:17-37.  The kernels are stateless across calls (v0 = 0,
    T folded in the launch), so reset() has nothing to clear.
    """

    def __init__(self, d_max: float = 8.0, norm: float = 8.0):
        super().__init__()
        self.d_max, self.norm = d_max, norm
        self.v = 0.0

    def reset(self):
        self.v = 0.0

    def extra_repr(self):
        return f"clamp=[0,{self.d_max}], out=level/{self.norm}"


class Holder(nn.Module):
    """Anonymous container; children are attached by dotted path."""


def attach(root: nn.Module, path: str, module: nn.Module) -> nn.Module:
    parts = path.split(".")
    cur = root
    for p in parts[:-1]:
        nxt = cur._modules.get(p)
        if nxt is None:
            nxt = Holder()
            cur.add_module(p, nxt)
        cur = nxt
    cur.add_module(parts[-1], module)
    return module


def conv2d(cin, cout, k, stride=1, pad=0, groups=1, bias=True):
    return nn.Conv2d(cin, cout, k, stride, pad, groups=groups, bias=bias)


def conv_bn(root, path, conv, bn_features, names=("0", "1"), bn1d=False):
    attach(root, f"{path}.{names[0]}", conv)
    attach(root, f"{path}.{names[1]}", (nn.BatchNorm1d if bn1d else nn.BatchNorm2d)(bn_features))


# ------------------------------------------------------------------ backbone (sdtv2.py)
def _rep_conv(root, path, c):
    """RepConv + outer BN: sdtv2.py:111-132 and :280-296."""
    attach(root, f"{path}.0.body.0", conv2d(c, c, 1, bias=False))
    attach(root, f"{path}.0.body.1.bn", nn.BatchNorm2d(c))
    attach(root, f"{path}.0.body.2.0", conv2d(c, c, 3, groups=c, bias=False))
    attach(root, f"{path}.0.body.2.1", conv2d(c, c, 1, bias=False))
    attach(root, f"{path}.0.body.2.2", nn.BatchNorm2d(c))
    attach(root, f"{path}.1", nn.BatchNorm2d(c))


def _downsample(root, path, cin, cout, k, stride, pad, first):
    """MS_DownSampling: sdtv2.py:386-421."""
    attach(root, f"{path}.encode_conv", conv2d(cin, cout, k, stride, pad))
    attach(root, f"{path}.encode_bn", nn.BatchNorm2d(cout))
    if not first:
        attach(root, f"{path}.encode_spike", NeuronSlot())


def _conv_block(root, path, dim, ratio):
    """MS_ConvBlock + SepConv: sdtv2.py:135-219."""
    med = 2 * dim
    attach(root, f"{path}.Conv.spike1", NeuronSlot())
    attach(root, f"{path}.Conv.pwconv1", conv2d(dim, med, 1, bias=False))
    attach(root, f"{path}.Conv.bn1", nn.BatchNorm2d(med))
    attach(root, f"{path}.Conv.spike2", NeuronSlot())
    attach(root, f"{path}.Conv.dwconv", conv2d(med, med, 7, pad=3, groups=med, bias=False))
    attach(root, f"{path}.Conv.pwconv2", conv2d(med, dim, 1, bias=False))
    attach(root, f"{path}.Conv.bn2", nn.BatchNorm2d(dim))
    attach(root, f"{path}.spike1", NeuronSlot())
    attach(root, f"{path}.conv1", conv2d(dim, dim * ratio, 3, pad=1, bias=False))
    attach(root, f"{path}.bn1", nn.BatchNorm2d(dim * ratio))
    attach(root, f"{path}.spike2", NeuronSlot())
    attach(root, f"{path}.conv2", conv2d(dim * ratio, dim, 3, pad=1, bias=False))
    attach(root, f"{path}.bn2", nn.BatchNorm2d(dim))


def _ms_block(root, path, dim, ratio):
    """MS_Block = SDSA + MS_MLP: sdtv2.py:222-383."""
    attach(root, f"{path}.attn.head_spike", NeuronSlot())
    for name in ("q_conv", "k_conv", "v_conv"):
        _rep_conv(root, f"{path}.attn.{name}", dim)
    for name in ("q_spike", "k_spike", "v_spike", "attn_spike"):
        attach(root, f"{path}.attn.{name}", NeuronSlot())
    _rep_conv(root, f"{path}.attn.proj_conv", dim)
    hid = int(dim * ratio)
    attach(root, f"{path}.mlp.fc1_conv", nn.Conv1d(dim, hid, 1))
    attach(root, f"{path}.mlp.fc1_bn", nn.BatchNorm1d(hid))
    attach(root, f"{path}.mlp.fc1_spike", NeuronSlot())
    attach(root, f"{path}.mlp.fc2_conv", nn.Conv1d(hid, dim, 1))
    attach(root, f"{path}.mlp.fc2_bn", nn.BatchNorm1d(dim))
    attach(root, f"{path}.mlp.fc2_spike", NeuronSlot())


def build_backbone_tree(root, in_channels, embed_dim, mlp_ratios):
    """Spiking_vit_MetaFormer.__init__: sdtv2.py:426-569."""
    e = embed_dim
    _downsample(root, "downsample1_1", in_channels, e[0] // 2, 7, 2, 3, True)
    _conv_block(root, "ConvBlock1_1.0", e[0] // 2, mlp_ratios)
    _downsample(root, "downsample1_2", e[0] // 2, e[0], 3, 2, 1, False)
    _conv_block(root, "ConvBlock1_2.0", e[0], mlp_ratios)
    _downsample(root, "downsample2", e[0], e[1], 3, 2, 1, False)
    _conv_block(root, "ConvBlock2_1.0", e[1], mlp_ratios)
    _conv_block(root, "ConvBlock2_2.0", e[1], mlp_ratios)
    _downsample(root, "downsample3", e[1], e[2], 3, 2, 1, False)
    for j in range(6):
        _ms_block(root, f"block3.{j}", e[2], mlp_ratios)
    _downsample(root, "downsample4", e[2], e[3], 3, 1, 1, False)
    for j in range(2):
        _ms_block(root, f"block4.{j}", e[3], mlp_ratios)


# ------------------------------------------------------------------ pixel decoder
def _sepconv_spike(root, path, dim, k, ratio=2):
    """SepConv_Spike: mmcv_spike/SNN_core.py:11-63."""
    med = int(ratio * dim)
    attach(root, f"{path}.spike1", NeuronSlot())
    conv_bn(root, f"{path}.pwconv1", conv2d(dim, med, 1, bias=False), med)
    attach(root, f"{path}.spike2", NeuronSlot())
    conv_bn(root, f"{path}.dwconv", conv2d(med, med, k, pad=(k - 1) // 2, groups=med, bias=False), med)
    attach(root, f"{path}.spike3", NeuronSlot())
    conv_bn(root, f"{path}.pwconv2", conv2d(med, dim, 1, bias=False), dim)


def _dcn(root, path, c, group, dw_k, ksz=3):
    """DCNv3_pytorch: ops_dcnv3/modules/dcnv3.py:96-196."""
    for n in ("dw_spike", "offset_spike", "mask_spike"):
        attach(root, f"{path}.{n}", NeuronSlot())
    conv_bn(root, f"{path}.dw_conv", conv2d(c, c, dw_k, pad=(dw_k - 1) // 2, groups=c, bias=False), c)
    conv_bn(root, f"{path}.offset", conv2d(c, group * ksz * ksz * 2, 1), group * ksz * ksz * 2)
    conv_bn(root, f"{path}.mask", conv2d(c, group * ksz * ksz, 1), group * ksz * ksz)
    _sepconv_spike(root, f"{path}.input_proj", c, dw_k)
    _sepconv_spike(root, f"{path}.output_proj", c, dw_k)
    # dcnv3.py:192-196 zero-initialises the offset / mask convolutions
    for n in ("offset", "mask"):
        conv = root.get_submodule(f"{path}.{n}.0")
        nn.init.zeros_(conv.weight); nn.init.zeros_(conv.bias)


def _enc_mlp(root, path, dim, hid):
    """MS_MLP (head version): mmcv_spike/transformer.py:787-831."""
    attach(root, f"{path}.fc1_spike", NeuronSlot())
    attach(root, f"{path}.fc1_conv", nn.Conv1d(dim, hid, 1))
    attach(root, f"{path}.fc1_bn", nn.BatchNorm1d(hid))
    attach(root, f"{path}.fc2_spike", NeuronSlot())
    attach(root, f"{path}.fc2_conv", nn.Conv1d(hid, dim, 1))
    attach(root, f"{path}.fc2_bn", nn.BatchNorm1d(dim))


def build_pixel_decoder_tree(root, in_channels, feat, out_channels, encoder):
    """DCNTransformerEncoderPixelDecoder.__init__: pixel_decoder.py:338-405 (+ parent :46-86)."""
    import torch

    n_in = len(in_channels)
    for i in range(n_in - 1):
        attach(root, f"lateral_convs.{i}.0", conv2d(in_channels[i], feat, 1))
        attach(root, f"lateral_convs.{i}.1", nn.BatchNorm2d(feat))
        attach(root, f"lateral_convs_spike.{i}", NeuronSlot())
        attach(root, f"output_convs.{i}.0", conv2d(feat, feat, 3, pad=1, groups=feat, bias=False))
        attach(root, f"output_convs.{i}.1", nn.BatchNorm2d(feat))
        attach(root, f"output_convs_spike.{i}", NeuronSlot())
    attach(root, "last_feat_conv_spike", NeuronSlot())
    attach(root, "mask_feature_spike", NeuronSlot())
    attach(root, "mask_feature", conv2d(feat, out_channels, 1))
    lc = encoder["layer_cfg"]
    sa, ffn = lc["self_attn_cfg"], lc["ffn_cfg"]
    assert sa["embed_dims"] == feat, "embed_dims of the encoder must equal feat_channels (pixel_decoder.py:388-391)"
    if sa["embed_dims"] % sa["group"] != 0:
        raise ValueError(f"channels must be divisible by group, but got {sa['embed_dims']} and {sa['group']}")
    for l in range(encoder["num_layers"]):
        p = f"encoder.layers.{l}"
        layer = attach(root, p, Holder())
        for g in ("gamma1", "gamma2", "gamma3"):  # detr_layers.py:301,329-332
            layer.register_parameter(g, nn.Parameter(1e-6 * torch.ones(feat)))
        _sepconv_spike(root, f"{p}.Conv", feat, 3)
        _dcn(root, f"{p}.dcn", feat, sa["group"], sa["dw_kernel_size"])
        _enc_mlp(root, f"{p}.ffn", ffn["embed_dims"], ffn["feedforward_channels"])
    attach(root, "encoder_in_proj_spike", NeuronSlot())  # built, never called: pixel_decoder.py:396 vs :435
    conv_bn(root, "encoder_in_proj", conv2d(in_channels[-1], feat, 1), feat)
    attach(root, "encoder_out_proj_spike", NeuronSlot())
    conv_bn(root, "encoder_out_proj", conv2d(feat, feat, 1), feat)


# ------------------------------------------------------------------ transformer decoder + head
def _attn_block(root, path, dim):
    """{Cross,}MultiHeadAttentionBlock: mmcv_spike/transformer.py:196-235, 280-316."""
    for n in ("q_conv", "k_conv", "v_conv", "out_conv"):
        conv_bn(root, f"{path}.{n}", nn.Conv1d(dim, dim, 1), dim, bn1d=True)
    for n in ("q_conv_spike", "k_conv_spike", "v_conv_spike", "q_spike", "k_spike", "v_spike", "attn_spike"):
        attach(root, f"{path}.{n}", NeuronSlot())


def build_head_tree(root, feat, out_channels, num_queries, num_classes, transformer_decoder):
    """mmdet MaskFormerHead.__init__: dense_heads/maskformer_head.py:68-168 (+ mmseg wrapper :36-51)."""
    import torch

    lc = transformer_decoder["layer_cfg"]
    dim = lc["self_attn_cfg"]["embed_dims"]
    hid = lc["ffn_cfg"]["feedforward_channels"]
    for l in range(transformer_decoder["num_layers"]):
        p = f"transformer_decoder.layers.{l}"
        _attn_block(root, f"{p}.self_attn.attn", dim)
        _attn_block(root, f"{p}.cross_attn.attn", dim)
        attach(root, f"{p}.ffn.fc1_spike", NeuronSlot())   # MSDA_FFN: transformer.py:710-766
        attach(root, f"{p}.ffn.fc1", nn.Conv1d(dim, hid, 1))
        attach(root, f"{p}.ffn.bn1", nn.BatchNorm1d(hid))
        attach(root, f"{p}.ffn.fc2_spike", NeuronSlot())
        attach(root, f"{p}.ffn.fc2", nn.Conv1d(hid, dim, 1))
        attach(root, f"{p}.ffn.bn2", nn.BatchNorm1d(dim))
    attach(root, "query_embed", nn.Embedding(num_queries, out_channels))
    attach(root, "query_feat", nn.Embedding(num_queries, out_channels))
    attach(root, "level_embed", nn.Embedding(3, feat))
    attach(root, "decoder_out_spike", NeuronSlot())
    attach(root, "cls_embed", nn.Linear(feat, num_classes + 1))
    attach(root, "mask_embed_spike", NeuronSlot())
    attach(root, "mask_embed.fc1", nn.Linear(feat, feat, bias=False))   # SNN_core.py:95-114
    attach(root, "mask_embed.spike1", NeuronSlot())
    attach(root, "mask_embed.fc2", nn.Linear(feat, feat, bias=False))
    attach(root, "mask_embed.spike2", NeuronSlot())
    attach(root, "mask_embed.fc_out", nn.Linear(feat, out_channels))
    root.register_parameter("w", nn.Parameter(torch.ones(1)))
    attach(root, "shortcut_conv_spike", NeuronSlot())
    conv_bn(root, "shortcut_conv", nn.Conv1d(num_queries, num_queries, 1, bias=False), num_queries, bn1d=True)
