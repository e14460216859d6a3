"""Host-side mirror of the reference's registered classes (the drop-in boundary).

Same class names, constructor keywords, parameter names and tensor contracts
as the reference:

  Spiking_vit_MetaFormer              Segmentation/mmseg/models/backbones/sdtv2.py:424-655
  DCNTransformerEncoderPixelDecoder   Segmentation/mmdet/models/layers/pixel_decoder.py:316-472
  MaskFormerHead                      Segmentation/mmseg/models/decode_heads/maskformer_head.py:22-180
                                      (+ mmdet parent dense_heads/maskformer_head.py:68-168, 498-586)
  EncoderDecoder (inference subset)   Segmentation/mmseg/models/segmentors/encoder_decoder.py:118-133

The modules own the parameters; the arithmetic is executed by engine.py on the
sm_100a kernels.  There is no CPU path: calling forward on CPU tensors raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import params
from .registry import MODELS, ConfigDict, register_everywhere, to_config


def _require_cuda(x: torch.Tensor, who: str):
    if not x.is_cuda:
        raise RuntimeError(f"{who}: CUDA tensor required -- spike2former_b200 has no CPU path")


class _Engined(nn.Module):
    """Shared plumbing: lazily built, invalidated whenever parameters are (re)loaded."""

    def __init__(self):
        super().__init__()
        self._plan = None

    def invalidate(self):
        self._plan = None

    def _load_from_state_dict(self, *a, **k):
        self._plan = None
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._plan = None
        return super()._apply(fn, *a, **k)

    def reset(self):
        """ResetModelHook protocol (resetmodel_hook.py:17-37): neurons are stateless here."""


@register_everywhere
class Spiking_vit_MetaFormer(_Engined):
    def __init__(self, img_size_h=128, img_size_w=128, patch_size=16, in_channels=2, num_classes=11,
                 embed_dim=(64, 128, 256), num_heads=(1, 2, 4), mlp_ratios=(4, 4, 4), qkv_bias=False, qk_scale=None,
                 drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, norm_layer=nn.LayerNorm, depths=(6, 8, 6),
                 sr_ratios=(8, 4, 2), T=1, decode_mode="snn", init_cfg=None,
                 norm_cfg=dict(type="BN", requires_grad=True), norm_eval=True, pretrained=None):
        super().__init__()
        if len(embed_dim) != 4:
            raise ValueError("embed_dim must list four stage widths (sdtv2.py:455-535)")
        if not isinstance(num_heads, int) or not isinstance(mlp_ratios, int):
            raise TypeError("num_heads and mlp_ratios are scalars in the SDTv2 configs (sdtv2.py:505-535)")
        for d in (embed_dim[2], embed_dim[3]):
            assert d % num_heads == 0, f"dim {d} should be divided by num_heads {num_heads}."  # sdtv2.py:271-273
        self.num_classes, self.depths, self.T = num_classes, depths, T
        self.decode_mode, self.norm_cfg, self.init_cfg = decode_mode, norm_cfg, init_cfg
        self.embed_dim, self.num_heads, self.mlp_ratios = list(embed_dim), num_heads, mlp_ratios
        self.in_channels = in_channels
        params.build_backbone_tree(self, in_channels, self.embed_dim, mlp_ratios)
        self.eval()

    def init_weights(self):
        """sdtv2.py:577-612: load `init_cfg.checkpoint`, strip 'backbone.', strict=False."""
        if self.init_cfg is None or "checkpoint" not in self.init_cfg:
            return
        ckpt = torch.load(self.init_cfg["checkpoint"], map_location="cpu")
        sd = ckpt.get("state_dict", ckpt.get("model", ckpt))
        sd = {(k[9:] if k.startswith("backbone.") else k): v for k, v in sd.items()}
        return self.load_state_dict(sd, strict=False)

    def forward(self, x):
        """fp32 [B,3,H,W] -> list of four maps.  'Qsnn': [T,B,C,H/2^i,W/2^i] (sdtv2.py:614-651)."""
        _require_cuda(x, type(self).__name__)
        from . import engine

        feats = engine.backbone_forward(self, x)
        return engine.export_backbone_feats(self, feats)


@register_everywhere
class DCNTransformerEncoderPixelDecoder(_Engined):
    def __init__(self, in_channels, feat_channels, out_channels, T=4, norm_cfg=dict(type="GN", num_groups=32),
                 act_cfg=dict(type="ReLU"), encoder=None, positional_encoding=dict(num_feats=128, normalize=True),
                 init_cfg=None):
        super().__init__()
        self.in_channels, self.feat_channels, self.out_channels = list(in_channels), feat_channels, out_channels
        self.num_inputs = len(in_channels)
        self.T = T  # stored, never used: pixel_decoder.py:365,434
        self.encoder_cfg = to_config(encoder)
        self.positional_encoding_cfg = positional_encoding
        params.build_pixel_decoder_tree(self, self.in_channels, feat_channels, out_channels, self.encoder_cfg)
        self.encoder_embed_dims = feat_channels
        self.eval()

    def init_weights(self):
        pass

    def forward(self, feats, batch_img_metas=None):
        """-> (mask_feature [T,B,C,H/2,W/2], memory, [y32, y64, y128]) as pixel_decoder.py:417-472."""
        from . import engine

        _require_cuda(feats[0], type(self).__name__)
        return engine.pixel_decoder_forward_public(self, feats)


@register_everywhere
class MaskFormerHead(_Engined):
    """mmseg wrapper + mmdet head in one class (the reference reaches it as `type='MaskFormerHead'`)."""

    def __init__(self, num_classes=150, align_corners=False, ignore_index=255, in_channels=None, feat_channels=256,
                 out_channels=256, num_queries=100, T=4, pixel_decoder=None, enforce_decoder_input_project=False,
                 transformer_decoder=None, positional_encoding=dict(num_feats=128, normalize=True), loss_cls=None,
                 loss_mask=None, loss_dice=None, train_cfg=None, test_cfg=None, init_cfg=None, **kwargs):
        super().__init__()
        if enforce_decoder_input_project or transformer_decoder["layer_cfg"]["self_attn_cfg"]["embed_dims"] != feat_channels:
            raise NotImplementedError("decoder_input_projs other than Identity are not used by any Spike2Former config")
        self.num_classes, self.align_corners, self.ignore_index = num_classes, align_corners, ignore_index
        self.out_channels = num_classes            # decode_heads/maskformer_head.py:46-48
        self.feat_channels, self.mask_channels = feat_channels, out_channels
        self.num_queries, self.T, self.alpha = num_queries, T, 4
        self.num_transformer_feat_level = 3
        self.transformer_decoder_cfg = to_config(transformer_decoder)
        self.num_transformer_decoder_layers = transformer_decoder["num_layers"]
        self.positional_encoding_cfg = to_config(positional_encoding)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        pd = dict(pixel_decoder)
        pd.update(in_channels=in_channels, feat_channels=feat_channels, out_channels=out_channels)
        self.pixel_decoder = MODELS.build(pd)
        params.build_head_tree(self, feat_channels, out_channels, num_queries, num_classes, self.transformer_decoder_cfg)
        self.eval()

    def init_weights(self):
        pass

    def forward(self, x, batch_data_samples=None):
        """-> (all_cls_scores [L,B,nq,K+1], all_mask_preds [L,B,nq,H/2,W/2]): dense_heads/maskformer_head.py:498-586."""
        from . import engine

        _require_cuda(x[0], type(self).__name__)
        return engine.head_forward_public(self, x)

    def predict(self, x, batch_img_metas, test_cfg=None):
        """-> seg logits fp32 [B,num_classes,H,W]: decode_heads/maskformer_head.py:138-180."""
        from . import engine

        _require_cuda(x[0], type(self).__name__)
        img_shape = tuple(batch_img_metas[0]["img_shape"])
        return engine.head_predict(self, x, img_shape)

    def loss(self, x, batch_data_samples, train_cfg=None):
        raise NotImplementedError("training glue (Hungarian matching + losses) is outside the hot path (SURVEY.md section 8f-3)")


@register_everywhere
class EncoderDecoder(_Engined):
    """Inference subset of mmseg's EncoderDecoder: extract_feat + decode_head.predict (whole mode)."""

    def __init__(self, backbone, decode_head, data_preprocessor=None, neck=None, auxiliary_head=None, train_cfg=None,
                 test_cfg=None, pretrained=None, init_cfg=None):
        super().__init__()
        self.backbone = MODELS.build(backbone)
        self.decode_head = MODELS.build(decode_head)
        self.test_cfg = test_cfg
        self.align_corners = self.decode_head.align_corners
        self.num_classes = self.decode_head.num_classes
        self.out_channels = self.decode_head.out_channels
        self._graphs = {}
        self.use_cuda_graph = True       # replay one captured CUDA graph per (input shape, output kind)

    def invalidate(self):
        super().invalidate()
        self._graphs = {}
        for m in (self.backbone, self.decode_head, self.decode_head.pixel_decoder):
            m.invalidate()

    def _load_from_state_dict(self, *a, **k):
        self._graphs = {}
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._graphs = {}
        return super()._apply(fn, *a, **k)

    def _run(self, inputs, labels):
        """The whole forward is ~300 kernel launches of 5-50 us each: it is captured once per input shape into a
        CUDA graph (all launches go to torch's current stream through the C ABI, workspaces come from the graph's
        private pool) and replayed.  The returned tensor is the graph's output buffer: it is overwritten by the next
        call with the same shape, exactly like a CUDA-graphed module in any serving stack."""
        from . import engine

        if not self.use_cuda_graph:
            return engine.segmentor_logits(self, inputs, labels=labels)
        if torch.cuda.is_current_stream_capturing():
            return engine.segmentor_logits(self, inputs, labels=labels)
        key = (tuple(inputs.shape), inputs.device.index, bool(labels))
        g = self._graphs.get(key)
        if g is None:
            g = engine.GraphedForward(self, inputs, labels)
            self._graphs[key] = g
        return g(inputs)

    def extract_feat(self, inputs):
        return self.backbone(inputs)

    def encode_decode(self, inputs, batch_img_metas=None):
        """fp32 [B,3,H,W] -> seg logits [B,K,H,W] (encoder_decoder.py:125-133), whole-image mode."""
        _require_cuda(inputs, type(self).__name__)
        return self._run(inputs, labels=False)

    def forward(self, inputs, data_samples=None, mode="tensor"):
        return self.encode_decode(inputs)

    @torch.no_grad()
    def predict_labels(self, inputs):
        """argmax over classes, as BaseSegmentor.postprocess_result (segmentors/base.py:177-188)."""
        _require_cuda(inputs, type(self).__name__)
        return self._run(inputs, labels=True)                          # argmax fused into the tail kernel (uint8)


def build_segmentor(cfg) -> EncoderDecoder:
    cfg = dict(cfg)
    cfg.pop("data_preprocessor", None)
    return MODELS.build(cfg)
