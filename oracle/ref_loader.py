"""Execute the reference's own hot-path files (TEST INFRASTRUCTURE ONLY).

The reference cannot be imported normally: mmengine / mmcv / timm /
spikingjelly are not installed and mmdet/models/layers/__init__.py:7 imports a
file that is missing from the tree.  Every hot-path file does
run when loaded *by path* after a handful of stand-in modules are registered
in sys.modules.  This module is that loader.  It is a loader, not a
restatement: the arithmetic executed is the reference's own bytes.

It only works where /root/reference exists (the build container).  The GPU box
never calls it; tests that need it skip when the tree is absent.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
import warnings

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("S2F_REFERENCE_ROOT", "/root/reference/Segmentation")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "mmseg", "models", "backbones"))


# --------------------------------------------------------------------------- stand-ins
class AttrDict(dict):
    """dict with attribute access (what the reference expects of mmengine.ConfigDict)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(key) from e

    __setattr__ = dict.__setitem__


def _attrify(obj):
    if isinstance(obj, dict):
        return AttrDict({k: _attrify(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return type(obj)(_attrify(v) for v in obj)
    return obj


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        pass


class _Registry:
    def __init__(self):
        self.table = {}

    def register_module(self, name=None, force=False, module=None):
        def wrap(cls):
            self.table[name or cls.__name__] = cls
            return cls

        return wrap

    def build(self, cfg, default_args=None):
        cfg = _attrify(dict(cfg))
        kind = cfg.pop("type").split(".")[-1]
        return self.table[kind](**cfg)


class _NoDropPath(nn.Identity):
    def __init__(self, p=0.0):
        super().__init__()


class _Sample:
    def __init__(self, metainfo=None):
        self.metainfo = metainfo or {}


class _NullLoss(nn.Module):
    def __init__(self, **kw):
        super().__init__()


def _passthrough_decorator(*a, **k):
    def wrap(f):
        return f

    return wrap


def _noop(*a, **k):
    return None


def _tn(t, std=0.02, **k):
    return nn.init.trunc_normal_(t, std=std)


_LOADED = None


def _shell(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    m.__dict__.update(attrs)
    if "." in name:
        parent, child = name.rsplit(".", 1)
        if parent in sys.modules:
            setattr(sys.modules[parent], child, m)
    return m


def _exec(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    parent, child = name.rsplit(".", 1)
    if parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    spec.loader.exec_module(m)
    return m


def load():
    """Load the reference hot-path files once; returns a namespace of handles."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    warnings.filterwarnings("ignore")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    models, task_utils = _Registry(), _Registry()
    two = lambda x: (x, x)
    _shell("timm"); _shell("timm.models")
    _shell("timm.models.layers", to_2tuple=two, trunc_normal_=_tn, DropPath=_NoDropPath)
    _shell("mmengine", ConfigDict=AttrDict)
    _shell("mmengine.model", BaseModule=_BaseModule, ModuleList=nn.ModuleList, Sequential=nn.Sequential,
           caffe2_xavier_init=_noop, constant_init=_noop, xavier_init=_noop)
    _shell("mmengine.model.weight_init", constant_init=_noop, trunc_normal_=_tn, trunc_normal_init=_noop)
    _shell("mmengine.logging", print_log=_noop)
    _shell("mmengine.runner", CheckpointLoader=None)
    _shell("mmengine.registry", MODELS=models)
    _shell("mmengine.utils", deprecated_api_warning=_passthrough_decorator, to_2tuple=two, digit_version=None)
    _shell("mmengine.config", ConfigDict=AttrDict)
    _shell("mmengine.structures", InstanceData=dict, PixelData=dict)
    _shell("mmcv")
    _shell("mmcv.cnn", Conv2d=nn.Conv2d, Linear=nn.Linear, ConvModule=None, build_activation_layer=None,
           build_conv_layer=None, build_norm_layer=None)
    _shell("mmcv.cnn.bricks"); _shell("mmcv.cnn.bricks.transformer", FFN=None)
    _shell("mmcv.ops", point_sample=None)
    _shell("spikingjelly"); _shell("spikingjelly.clock_driven")
    _shell("spikingjelly.clock_driven.neuron", MultiStepParametricLIFNode=None, MultiStepLIFNode=None)
    _shell("mmseg"); _shell("mmseg.registry", MODELS=models)
    _shell("mmseg.models"); _shell("mmseg.models.utils"); _shell("mmseg.structures")
    _shell("mmseg.structures.seg_data_sample", SegDataSample=_Sample)
    _shell("mmseg.utils", ConfigType=dict, SampleList=list)
    _shell("mmdet"); _shell("mmdet.registry", MODELS=models, TASK_UTILS=task_utils)
    _shell("mmdet.utils", ConfigType=dict, OptConfigType=dict, OptMultiConfig=dict, MultiConfig=dict,
           InstanceList=list, reduce_mean=None)
    _shell("mmdet.structures", SampleList=list)
    _shell("mmdet.models")
    _shell("mmdet.models.utils", get_uncertain_point_coords_with_randomness=None, multi_apply=None,
           preprocess_panoptic_gt=None)
    pk = "mmdet.models.layers"
    for sub in ("", ".transformer", ".transformer.mmcv_spike", ".transformer.ops_dcnv3",
                ".transformer.ops_dcnv3.modules", ".transformer.ops_dcnv3.functions"):
        _shell(pk + sub)
    _shell("mmdet.models.dense_heads")

    ns = types.SimpleNamespace(MODELS=models)
    _exec("mmseg.models.utils.Qtrick", "mmseg/models/utils/Qtrick.py")
    _exec("mmdet.models.utils.Qtrick", "mmdet/models/utils/Qtrick.py")
    ns.sdtv2 = _exec("mmseg.models.backbones_sdtv2", "mmseg/models/backbones/sdtv2.py")
    ns.snn_core = _exec(pk + ".transformer.mmcv_spike.SNN_core", "mmdet/models/layers/transformer/mmcv_spike/SNN_core.py")
    ns.dcn_func = _exec(pk + ".transformer.ops_dcnv3.functions.dcnv3_func",
                        "mmdet/models/layers/transformer/ops_dcnv3/functions/dcnv3_func.py")
    sys.modules[pk + ".transformer.ops_dcnv3.functions"].dcnv3_core_pytorch = ns.dcn_func.dcnv3_core_pytorch
    ns.dcn_mod = _exec(pk + ".transformer.ops_dcnv3.modules.dcnv3",
                       "mmdet/models/layers/transformer/ops_dcnv3/modules/dcnv3.py")
    ns.spike_tr = _exec(pk + ".transformer.mmcv_spike.transformer",
                        "mmdet/models/layers/transformer/mmcv_spike/transformer.py")
    ns.detr = _exec(pk + ".transformer.detr_layers", "mmdet/models/layers/transformer/detr_layers.py")
    tpk = sys.modules[pk + ".transformer"]
    tpk.DetrTransformerEncoder = ns.detr.DetrTransformerEncoder
    tpk.DCNDetrTransformerEncoder = ns.detr.DCNDetrTransformerEncoder
    ns.pe = _exec(pk + ".positional_encoding", "mmdet/models/layers/positional_encoding.py")
    ns.pixel_decoder = _exec(pk + ".pixel_decoder", "mmdet/models/layers/pixel_decoder.py")
    lpk = sys.modules[pk]
    lpk.DetrTransformerDecoder = ns.detr.DetrTransformerDecoder
    lpk.SinePositionalEncoding = ns.pe.SinePositionalEncoding
    _shell(pk + ".transformer.utils", QueryProposal=None)

    class AnchorFreeHead(_BaseModule):
        pass

    _shell("mmdet.models.dense_heads.anchor_free_head", AnchorFreeHead=AnchorFreeHead)
    for lname in ("CrossEntropyLoss", "FocalLoss", "DiceLoss"):
        models.table[lname] = type(lname, (_NullLoss,), {})
    ns.det_head = _exec("mmdet.models.dense_heads.maskformer_head", "mmdet/models/dense_heads/maskformer_head.py")
    sys.modules["mmdet.models.dense_heads"].MaskFormerHead = ns.det_head.MaskFormerHead
    ns.seg_head = _exec("mmseg.models.decode_heads_maskformer_head", "mmseg/models/decode_heads/maskformer_head.py")
    from Qtrick_architecture.clock_driven import neuron as _neuron, surrogate as _surrogate

    ns.neuron, ns.surrogate = _neuron, _surrogate
    _LOADED = ns
    return ns


# --------------------------------------------------------------------------- model handles
def build_reference(cfg):
    """cfg: the dict made by oracle.configs (backbone / decode_head literal dicts).

    Returns (backbone, head) in eval mode, built by the reference's registry.
    """
    import copy

    ns = load()
    cfg = copy.deepcopy(cfg)
    bb_cfg = dict(cfg["backbone"]); bb_cfg.pop("type"); bb_cfg.pop("init_cfg", None)
    backbone = ns.sdtv2.Spiking_vit_MetaFormer(**bb_cfg)
    head_cfg = dict(cfg["decode_head"]); head_cfg["train_cfg"] = None
    head = ns.MODELS.build(head_cfg)
    backbone.eval(); head.eval()
    return backbone, head


def reset_neurons(*modules):
    """What ResetModelHook -> spikingjelly functional.reset_net does (resetmodel_hook.py:17-37)."""
    for mod in modules:
        for m in mod.modules():
            if hasattr(m, "reset"):
                m.reset()


@torch.no_grad()
def reference_predict(backbone, head, img):
    """Reference inference: reset -> backbone -> head.predict (encoder_decoder.py:125-133)."""
    reset_neurons(backbone, head)
    b, _, h, w = img.shape
    metas = [dict(img_shape=(h, w), ori_shape=(h, w), pad_shape=(h, w)) for _ in range(b)]
    return head.predict(backbone(img), metas, dict(mode="whole"))


class SpikeTap:
    """Forward hooks on every Q_IFNode: records (name, pre-activation, output) per call."""

    def __init__(self, backbone, head, keep_tensors=True):
        ns = load()
        self.records = []
        self.handles = []
        for prefix, mod in (("backbone", backbone), ("decode_head", head)):
            for name, m in mod.named_modules():
                if isinstance(m, ns.neuron.Q_IFNode):
                    self.handles.append(m.register_forward_hook(self._hook(prefix + "." + name, keep_tensors)))

    def _hook(self, name, keep):
        def fn(mod, inp, out):
            x = inp[0].detach()
            self.records.append((name, x.clone() if keep else tuple(x.shape), out.detach().clone() if keep else None))

        return fn

    def close(self):
        for h in self.handles:
            h.remove()


# --------------------------------------------------------------------------- data preprocessor
class _BaseDataPreprocessor(nn.Module):
    """mmengine.model.BaseDataPreprocessor stand-in: cast_data is the identity on CPU."""

    def cast_data(self, data):
        return data


def load_data_preprocessor():
    """Execute mmseg/utils/misc.py (stack_batch) and mmseg/models/data_preprocessor.py; returns the reference class."""
    load()
    sys.modules["mmengine.model"].BaseDataPreprocessor = _BaseDataPreprocessor
    _shell("mmseg.utils.typing_utils", SampleList=list)       # misc.py:8 `from .typing_utils import SampleList`
    misc = _exec("mmseg.utils.misc", "mmseg/utils/misc.py")
    sys.modules["mmseg.utils"].stack_batch = misc.stack_batch
    mod = _exec("mmseg.models.data_preprocessor_ref", "mmseg/models/data_preprocessor.py")
    return mod.SegDataPreProcessor


# --------------------------------------------------------------------------- training glue
class _InstanceData:
    """mmengine.structures.InstanceData stand-in: attribute bag whose len() is the length of its first field."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __len__(self):
        for v in self.__dict__.values():
            return len(v)
        return 0


class _Field:
    def __init__(self, data):
        self.data = data

    @property
    def shape(self):
        return self.data.shape


class TrainSample:
    """SegDataSample stand-in for the loss path: `.gt_sem_seg.data`, `.metainfo`, `.set_metainfo`."""

    def __init__(self, gt_sem_seg, img_shape):
        self.gt_sem_seg = _Field(gt_sem_seg)
        self.metainfo = dict(img_shape=tuple(img_shape), ori_shape=tuple(img_shape), pad_shape=tuple(img_shape))

    def set_metainfo(self, d):
        self.metainfo.update(d)


def _multi_apply(func, *args, **kwargs):
    from functools import partial

    pfunc = partial(func, **kwargs) if kwargs else func
    return tuple(map(list, zip(*map(pfunc, *args))))


_TRAIN = None


def load_training():
    """Execute the reference's OWN loss / matching code: mmdet/models/losses/{utils,cross_entropy_loss,focal_loss,
    dice_loss}.py, task_modules/assigners/{assign_result,base_assigner,match_cost,hungarian_assigner}.py and
    task_modules/samplers/{sampling_result,mask_sampling_result,base_sampler,mask_pseudo_sampler}.py, and point the
    already loaded heads at the real InstanceData / multi_apply / reduce_mean semantics."""
    global _TRAIN
    if _TRAIN is not None:
        return _TRAIN
    ns = load()
    models = ns.MODELS
    task_utils = sys.modules["mmdet.registry"].TASK_UTILS

    class NiceRepr:
        pass

    _shell("mmdet.utils.util_mixins", NiceRepr=NiceRepr)
    sys.modules["mmdet.utils"].util_mixins = sys.modules["mmdet.utils.util_mixins"]
    _shell("mmdet.utils.util_random", ensure_rng=lambda rng=None: rng)
    _shell("mmdet.structures.bbox", bbox_overlaps=None, bbox_xyxy_to_cxcywh=None, BaseBoxes=type("BaseBoxes", (), {}),
           cat_boxes=None)
    sys.modules["mmengine.structures"].InstanceData = _InstanceData
    sys.modules["mmcv.ops"].sigmoid_focal_loss = None                     # CUDA-only op; the CPU path never calls it
    _shell("mmdet.models.losses")
    _exec("mmdet.models.losses.utils", "mmdet/models/losses/utils.py")
    for f in ("cross_entropy_loss", "focal_loss", "dice_loss"):            # register the REAL losses over the null ones
        _exec("mmdet.models.losses." + f, f"mmdet/models/losses/{f}.py")
    _shell("mmdet.models.task_modules")
    ap = "mmdet.models.task_modules.assigners"
    _shell(ap)
    ar = _exec(ap + ".assign_result", "mmdet/models/task_modules/assigners/assign_result.py")
    sys.modules[ap].AssignResult = ar.AssignResult
    _exec(ap + ".base_assigner", "mmdet/models/task_modules/assigners/base_assigner.py")
    _exec(ap + ".match_cost", "mmdet/models/task_modules/assigners/match_cost.py")
    _exec(ap + ".hungarian_assigner", "mmdet/models/task_modules/assigners/hungarian_assigner.py")
    sp = "mmdet.models.task_modules.samplers"
    _shell(sp)
    _exec(sp + ".sampling_result", "mmdet/models/task_modules/samplers/sampling_result.py")
    _exec(sp + ".mask_sampling_result", "mmdet/models/task_modules/samplers/mask_sampling_result.py")
    _exec(sp + ".base_sampler", "mmdet/models/task_modules/samplers/base_sampler.py")
    _exec(sp + ".mask_pseudo_sampler", "mmdet/models/task_modules/samplers/mask_pseudo_sampler.py")
    # names the head modules bound at import time
    for mod in (ns.det_head, ns.seg_head):
        mod.InstanceData = _InstanceData
    ns.det_head.multi_apply = _multi_apply
    ns.det_head.reduce_mean = lambda t: t                                  # single process (dist_utils.py:59-62)
    ns.task_utils = task_utils
    assert "HungarianAssigner" in task_utils.table and "FocalLoss" in models.table
    _TRAIN = ns
    return ns


def build_reference_for_training(cfg, train_cfg):
    """(backbone, head) with the reference's real losses, HungarianAssigner and MaskPseudoSampler, in train mode."""
    import copy

    ns = load_training()
    cfg = copy.deepcopy(cfg)
    bb_cfg = dict(cfg["backbone"]); bb_cfg.pop("type"); bb_cfg.pop("init_cfg", None)
    backbone = ns.sdtv2.Spiking_vit_MetaFormer(**bb_cfg)
    head_cfg = dict(cfg["decode_head"]); head_cfg["train_cfg"] = copy.deepcopy(train_cfg)
    head = ns.MODELS.build(head_cfg)
    backbone.train(); head.train()
    return backbone, head


def reference_train_loss(backbone, head, img, gt_sem_seg):
    """EncoderDecoder.loss (encoder_decoder.py:163-188) after ResetModelHook: dict of the reference's loss terms."""
    reset_neurons(backbone, head)
    samples = [TrainSample(gt_sem_seg[i], img.shape[-2:]) for i in range(img.shape[0])]      # [1, H, W] each, as PackSegInputs
    return head.loss(backbone(img), samples, None)
