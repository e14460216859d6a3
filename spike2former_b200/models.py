"""Host-side mirror of the reference's registered classes (the drop-in boundary).

Same class names, constructor keywords, parameter names and tensor contracts
as the reference:

  Spiking_vit_MetaFormer              Segmentation/mmseg/models/backbones/sdtv2.py:424-655
  DCNTransformerEncoderPixelDecoder   Segmentation/mmdet/models/layers/pixel_decoder.py:316-472
  MaskFormerHead                      Segmentation/mmseg/models/decode_heads/maskformer_head.py:22-180
                                      (+ mmdet parent dense_heads/maskformer_head.py:68-168, 498-586)
  EncoderDecoder (inference subset)   Segmentation/mmseg/models/segmentors/encoder_decoder.py:118-133
  SegDataPreProcessor                 Segmentation/mmseg/models/data_preprocessor.py:14-152 (+ utils/misc.py:30-110)

The modules own the parameters; the arithmetic is executed by engine.py on the
sm_100a kernels.  There is no CPU path: calling forward on CPU tensors raises.
"""
from __future__ import annotations

import collections

import torch
import torch.nn as nn

from . import params
from .registry import MODELS, ConfigDict, register_everywhere, to_config


def _require_cuda(x: torch.Tensor, who: str):
    if not x.is_cuda:
        raise RuntimeError(f"{who}: CUDA tensor required -- spike2former_b200 has no CPU path")


class _Engined(nn.Module):
    """Shared plumbing: the folded / packed device copy of the parameters (`_plan`) is built lazily and dropped
    whenever the parameters change: state_dict loads, `.to()` / `.cuda()`, `invalidate()`, and in-place edits of any
    parameter or buffer (detected through the tensors' version counters, see `weights_version`; writes through
    `param.data` bypass those counters by PyTorch's design -- call `invalidate()` after such an edit)."""

    def __init__(self):
        super().__init__()
        self._plan = None
        self._plan_version = None
        self._epoch = 0            # bumped by every explicit invalidation

    def invalidate(self):
        self._plan = None
        self._epoch += 1

    def _load_from_state_dict(self, *a, **k):
        self.invalidate()
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)

    def weights_version(self):
        """Changes whenever a parameter / buffer of this module (children included) is replaced or edited in place."""
        ts = self.__dict__.get("_tensors")
        if ts is None or ts[0] != self._epoch:
            ts = (self._epoch, [t for t in list(self.parameters()) + list(self.buffers())])
            self.__dict__["_tensors"] = ts
        return (self._epoch, sum(t._version for t in ts[1]), sum(m._epoch for m in self.modules() if isinstance(m, _Engined)))

    def current_plan(self, kind, device):
        v = self.weights_version()
        if self._plan is None or self._plan_version != v:
            self._plan, self._plan_version = kind(self, device), v
        return self._plan

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
        """-> dict of the 21 loss terms (decode_heads/maskformer_head.py:108-136, dense_heads/maskformer_head.py:367-496).
        x: the backbone's four maps ([T,B,C,H,W] or [T*B,C,H,W]) with autograd history; see spike2former_b200/train.py."""
        from . import train

        _require_cuda(x[0], type(self).__name__)
        owner = getattr(self, "_segmentor", None)
        if owner is None:
            raise RuntimeError("MaskFormerHead.loss needs the EncoderDecoder that owns this head (parameter names are rooted there)")
        net = owner()._training_net()
        feats = [t.flatten(0, 1) if t.dim() == 5 else t for t in x]
        gt = torch.stack([ds.gt_sem_seg.data for ds in batch_data_samples]).to(feats[0].device)
        return train.head_losses(net, feats, gt, self.ignore_index)


@register_everywhere
class SegDataPreProcessor(nn.Module):
    """uint8 images -> normalised fp32 batch, on the device (data_preprocessor.py:14-152).

    Same constructor keywords as the reference.  `forward(data, training)` takes `data['inputs']` as a list of uint8
    [3,H,W] tensors (what PackSegInputs produces) or an already stacked uint8 [B,3,H,W] / [B,H,W,3] batch and returns
    `dict(inputs=fp32 [B,3,Hp,Wp], data_samples=...)`.  The returned batch is an NCHW *view* of channels-last memory --
    the layout the stem kernel reads -- so the backbone consumes it without a copy."""

    def __init__(self, mean=None, std=None, size=None, size_divisor=None, pad_val=0, seg_pad_val=255, bgr_to_rgb=False,
                 rgb_to_bgr=False, batch_augments=None, test_cfg=None):
        super().__init__()
        assert not (bgr_to_rgb and rgb_to_bgr), "`bgr2rgb` and `rgb2bgr` cannot be set to True at the same time"
        self.size, self.size_divisor, self.pad_val, self.seg_pad_val = size, size_divisor, pad_val, seg_pad_val
        self.channel_conversion = rgb_to_bgr or bgr_to_rgb
        if mean is not None:
            assert std is not None, ("To enable the normalization in preprocessing, please specify both `mean` and `std`.")
            self._enable_normalize = True
            self.register_buffer("mean", torch.tensor(mean).view(-1, 1, 1), False)
            self.register_buffer("std", torch.tensor(std).view(-1, 1, 1), False)
            # host copies of the fp32 buffer values: the kernel takes them by value (no device read during graph capture)
            self._mean_host, self._std_host = self.mean.flatten().tolist(), self.std.flatten().tolist()
        else:
            self._enable_normalize = False
        if batch_augments is not None:
            raise NotImplementedError("batch_augments are training-time host glue (SURVEY.md section 2: OUT)")
        self.batch_augments = None
        self.test_cfg = test_cfg

    def _padded_size(self, H, W, size, size_divisor):
        if size is not None:
            assert size_divisor is None, "only one of size and size_divisor should be valid"
            return max(H, int(size[-2])), max(W, int(size[-1]))
        if size_divisor is not None and size_divisor > 1:
            return (H + size_divisor - 1) // size_divisor * size_divisor, (W + size_divisor - 1) // size_divisor * size_divisor
        return H, W

    def normalized(self, batch_u8, size=None, size_divisor=None, out=None):
        """stacked uint8 batch -> fp32 channels-last memory [B,Hp,Wp,3] (one kernel launch)."""
        from . import ops

        _require_cuda(batch_u8, type(self).__name__)
        chw = batch_u8.shape[1] == 3 and batch_u8.shape[3] != 3
        H, W = (batch_u8.shape[2], batch_u8.shape[3]) if chw else (batch_u8.shape[1], batch_u8.shape[2])
        hp, wp = self._padded_size(int(H), int(W), size, size_divisor)
        mean = self._mean_host if self._enable_normalize else None
        std = self._std_host if self._enable_normalize else None
        return ops.preprocess_u8(batch_u8, mean=mean, std=std, swap_rb=self.channel_conversion, size=(hp, wp),
                                 pad_val=float(self.pad_val), out=out)

    def forward(self, data, training=False):
        inputs = data["inputs"]
        data_samples = data.get("data_samples", None)
        if isinstance(inputs, (list, tuple)):
            img_size = inputs[0].shape[1:]
            assert all(i.shape[1:] == img_size for i in inputs), "The image size in a batch should be the same."
            inputs = torch.stack([i.cuda(non_blocking=True) for i in inputs], dim=0)
        else:
            inputs = inputs.cuda(non_blocking=True)
        size, div = (self.size, self.size_divisor) if training else \
            ((self.test_cfg.get("size", None), self.test_cfg.get("size_divisor", None)) if self.test_cfg else (None, None))
        if training:
            assert data_samples is not None, "During training, `data_samples` must be define."
        if training or self.test_cfg:
            assert (size is not None) ^ (div is not None), "only one of size and size_divisor should be valid"   # misc.py:63-64
        x = self.normalized(inputs, size, div)
        H0, W0 = int(img_h(inputs)), int(img_w(inputs))
        padding_size = (0, int(x.shape[2]) - W0, 0, int(x.shape[1]) - H0)    # (left, right, top, bottom): misc.py:78-86
        if training:
            # stack_batch with data_samples (misc.py:94-111): pad the label maps right / bottom with seg_pad_val and record
            # the UNPADDED image shape, the padded label shape and the padding
            for ds in data_samples:
                for field in ("gt_sem_seg", "gt_edge_map"):
                    if field in ds if hasattr(ds, "__contains__") else hasattr(ds, field):
                        f_ = getattr(ds, field)
                        f_.data = nn.functional.pad(f_.data, padding_size, value=self.seg_pad_val)
                ds.set_metainfo({"img_shape": torch.Size((H0, W0)), "pad_shape": ds.gt_sem_seg.shape,
                                 "padding_size": padding_size})
        elif self.test_cfg and data_samples is not None:
            # test-time padding (data_preprocessor.py:140-150, misc.py:112-116): only these two keys; img_shape is untouched
            for ds in data_samples:
                ds.set_metainfo({"img_padding_size": padding_size, "pad_shape": torch.Size((int(x.shape[1]), int(x.shape[2])))})
        return dict(inputs=x.permute(0, 3, 1, 2), data_samples=data_samples)


def img_h(b):
    return b.shape[2] if (b.shape[1] == 3 and b.shape[3] != 3) else b.shape[1]


def img_w(b):
    return b.shape[3] if (b.shape[1] == 3 and b.shape[3] != 3) else b.shape[2]


@register_everywhere
class EncoderDecoder(_Engined):
    """Inference subset of mmseg's EncoderDecoder: extract_feat + decode_head.predict (whole mode)."""

    def __init__(self, backbone, decode_head, data_preprocessor=None, neck=None, auxiliary_head=None, train_cfg=None,
                 test_cfg=None, pretrained=None, init_cfg=None):
        super().__init__()
        self.backbone = MODELS.build(backbone)
        self.decode_head = MODELS.build(decode_head)
        self.data_preprocessor = MODELS.build(data_preprocessor) if isinstance(data_preprocessor, dict) else data_preprocessor
        self.test_cfg = test_cfg
        import weakref

        self.decode_head._segmentor = weakref.ref(self)
        self._train_net = None
        self.align_corners = self.decode_head.align_corners
        self.num_classes = self.decode_head.num_classes
        self.out_channels = self.decode_head.out_channels
        self._graphs = collections.OrderedDict()
        self._shape_hits = collections.OrderedDict()
        self.use_cuda_graph = True       # replay one captured CUDA graph per (input shape, output kind)
        self.graph_cache_size = 4        # LRU bound: every cached graph owns a private activation pool (GBs at batch 32)
        self.graph_min_hits = 2          # a shape is captured the second time it is seen; one-off shapes run eagerly
        self.alias_graph_output = False  # True: return the graph's static output buffer (overwritten by the next call
                                         # with the same shape) instead of a fresh copy -- for serving loops that consume
                                         # the result before the next call

    def invalidate(self):
        super().invalidate()
        self._graphs.clear()
        for m in (self.backbone, self.decode_head, self.decode_head.pixel_decoder):
            m.invalidate()

    def _run(self, inputs, labels):
        """The whole forward is ~300 kernel launches of 5-50 us each: a shape that recurs is captured into a CUDA
        graph (all launches go to torch's current stream through the C ABI, workspaces come from the graph's private
        pool) and replayed.  A graph holds raw device pointers of the plans it was captured with, so it is keyed on
        the weights version of the whole model and holds references to those plans: any parameter change (child
        `load_state_dict`, `.to()`, in-place edits) makes the next call re-capture instead of replaying stale or
        freed memory.  The cache is a small LRU, since evaluation over variable-size images would otherwise keep one
        multi-GB activation pool per shape."""
        from . import engine

        if not self.use_cuda_graph or torch.cuda.is_current_stream_capturing():
            return engine.segmentor_logits(self, inputs, labels=labels)
        version = self.weights_version()
        key = (tuple(inputs.shape), inputs.device.index, bool(labels), inputs.dtype)
        g = self._graphs.get(key)
        if g is not None and g.version != version:
            del self._graphs[key]
            g = None
        if g is None:
            hits = self._shape_hits.pop(key, 0) + 1
            self._shape_hits[key] = hits
            while len(self._shape_hits) > 64:
                self._shape_hits.popitem(last=False)
            if hits < self.graph_min_hits:
                return engine.segmentor_logits(self, inputs, labels=labels)
            while len(self._graphs) >= max(1, self.graph_cache_size):
                self._graphs.popitem(last=False)             # frees the evicted graph and its pool
            g = engine.GraphedForward(self, inputs, labels)
            g.version = self.weights_version()               # the capture built the plans: same version afterwards
            self._graphs[key] = g
        else:
            self._graphs.move_to_end(key)
        out = g(inputs)
        return out if self.alias_graph_output else out.clone()

    def extract_feat(self, inputs):
        return self.backbone(inputs)

    def _training_net(self):
        from . import train

        if self._train_net is None or self._train_net[0] != self._epoch:
            self._train_net = (self._epoch, train.Net(self))
        return self._train_net[1]

    def loss(self, inputs, data_samples):
        """EncoderDecoder.loss (encoder_decoder.py:163-188): fp32 [B,3,H,W] + SegDataSamples (`gt_sem_seg.data` [1,H,W])
        -> dict of loss terms prefixed 'decode.' (add_prefix, :156-161), with autograd history on the parameters.
        Surrogate-gradient training mode: batch-statistics BatchNorm, see spike2former_b200/train.py."""
        from . import train

        _require_cuda(inputs, type(self).__name__)
        net = self._training_net()
        gt = torch.stack([ds.gt_sem_seg.data for ds in data_samples]).to(inputs.device)
        losses = train.head_losses(net, net.backbone(inputs), gt, self.decode_head.ignore_index)
        return {"decode." + k: v for k, v in losses.items()}

    def encode_decode(self, inputs, batch_img_metas=None):
        """fp32 [B,3,H,W] -> seg logits [B,K,H,W] (encoder_decoder.py:125-133), whole-image mode.
        A uint8 batch ([B,3,H,W] or [B,H,W,3]) is first normalised by `self.data_preprocessor` inside the same
        captured graph (BaseSegmentor.test_step = data_preprocessor + predict, mmengine base_model)."""
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
    dp = cfg.get("data_preprocessor", None)
    if isinstance(dp, dict):
        dp = dict(dp)
        dp.setdefault("type", "SegDataPreProcessor")
        cfg["data_preprocessor"] = dp
    return MODELS.build(cfg)
