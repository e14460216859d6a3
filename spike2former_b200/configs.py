"""Model dictionaries with the reference's config keys.

These are the `model=dict(backbone=..., decode_head=...)` literals of
Segmentation/configs/Spike2Former/SDTv2_maskformer_DCNpixelDecoder_ade20k.py:24-131
and ..._DCNPixelDecoder_CityScapes.py (same keys; num_classes=19, encoder
feedforward_channels=2048), so a reference config drops into
`spike2former_b200.registry.MODELS.build` unchanged.  `tiny()` is a reduced
width/depth variant with the same structure, used for fast parity tests.
"""
from __future__ import annotations

import copy


def _model(embed_dim4, feat, num_classes, num_queries, enc_ffn, dec_ffn, num_heads, group, img, enc_layers=6,
           dec_layers=6, crop=None, pre_test_cfg=None):
    norm_cfg = dict(type="SyncBN", requires_grad=True)
    ps_dim = feat // 2
    return dict(
        type="EncoderDecoder",
        # cfg:13-20: ImageNet statistics, BGR -> RGB; ADE20K: crop 512 x 512, no test-time padding; Cityscapes
        # (..._CityScapes.py:12-20): crop 512 x 1024 and test_cfg=dict(size_divisor=32)
        data_preprocessor=dict(type="SegDataPreProcessor", size=crop or (img, img), mean=[123.675, 116.28, 103.53],
                               std=[58.395, 57.12, 57.375], bgr_to_rgb=True, pad_val=0, seg_pad_val=255,
                               **({"test_cfg": pre_test_cfg} if pre_test_cfg else {})),
        backbone=dict(
            type="Spiking_vit_MetaFormer", img_size_h=img, img_size_w=img, patch_size=16, embed_dim=list(embed_dim4),
            num_heads=num_heads, mlp_ratios=4, in_channels=3, num_classes=num_classes, qkv_bias=False, depths=8,
            sr_ratios=1, T=1, norm_eval=True, norm_cfg=norm_cfg, decode_mode="Qsnn"),
        decode_head=dict(
            type="MaskFormerHead",
            in_channels=[embed_dim4[0] // 2, embed_dim4[0], embed_dim4[1], embed_dim4[3]],
            feat_channels=feat, in_index=[0, 1, 2, 3], num_classes=num_classes, out_channels=feat,
            num_queries=num_queries,
            pixel_decoder=dict(
                type="mmdet.DCNTransformerEncoderPixelDecoder", norm_cfg=norm_cfg, T=4,
                encoder=dict(num_layers=enc_layers, layer_cfg=dict(
                    self_attn_cfg=dict(embed_dims=feat, num_heads=num_heads, batch_first=True, dw_kernel_size=5,
                                       group=group),
                    ffn_cfg=dict(embed_dims=feat, feedforward_channels=enc_ffn, num_fcs=2))),
                positional_encoding=dict(num_feats=ps_dim, normalize=True)),
            enforce_decoder_input_project=False,
            positional_encoding=dict(num_feats=ps_dim, normalize=True),
            transformer_decoder=dict(
                return_intermediate=True, num_layers=dec_layers,
                layer_cfg=dict(
                    self_attn_cfg=dict(embed_dims=feat, num_heads=num_heads, attn_type="SA", batch_first=True),
                    cross_attn_cfg=dict(embed_dims=feat, num_heads=num_heads, attn_type="CA", batch_first=True),
                    ffn_cfg=dict(embed_dims=feat, feedforward_channels=dec_ffn, num_fcs=2, add_identity=True)),
                init_cfg=None),
            loss_cls=dict(type="mmdet.CrossEntropyLoss", use_sigmoid=False, loss_weight=1.0, reduction="mean",
                          class_weight=[1.0] * num_classes + [0.1]),
            loss_mask=dict(type="mmdet.FocalLoss", use_sigmoid=True, gamma=2.0, alpha=0.25, reduction="mean",
                           loss_weight=20.0),
            loss_dice=dict(type="mmdet.DiceLoss", use_sigmoid=True, activate=True, reduction="mean", naive_dice=True,
                           eps=1.0, loss_weight=1.0),
            train_cfg=None),
        train_cfg=dict(), test_cfg=dict(mode="whole"))


def ade20k():
    """SDTv2 + DCN pixel decoder, ADE20K (150 classes, 100 queries, 512x512)."""
    return _model([64, 128, 256, 360], 256, 150, 100, 1024, 2048, 8, 32, 512)


def cityscapes():
    """SDTv2 + DCN pixel decoder, Cityscapes (19 classes, encoder FFN 2048)."""
    return _model([64, 128, 256, 360], 256, 19, 100, 2048, 2048, 8, 32, 512, crop=(512, 1024),
                  pre_test_cfg=dict(size_divisor=32))


def tiny():
    """Same topology, narrow: used by the fast CPU/GPU parity tests and the committed goldens."""
    return _model([32, 32, 64, 96], 64, 11, 20, 128, 128, 4, 8, 64, enc_layers=2, dec_layers=6)


def clone(cfg):
    return copy.deepcopy(cfg)
