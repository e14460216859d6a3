"""CPU: drop-in boundary (registry, constructor keys, state_dict names), host-side folding, C ABI exports."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn.functional as F

import spike2former_b200 as s2f
from oracle import port
from spike2former_b200 import _lib, fold, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_registry_builds_reference_config_names():
    cfg = s2f.configs.ade20k()
    seg = s2f.MODELS.build({k: v for k, v in cfg.items() if k != "data_preprocessor"})
    assert type(seg.backbone).__name__ == "Spiking_vit_MetaFormer"
    assert type(seg.decode_head).__name__ == "MaskFormerHead"
    assert type(seg.decode_head.pixel_decoder).__name__ == "DCNTransformerEncoderPixelDecoder"
    assert "mmdet.DCNTransformerEncoderPixelDecoder" in s2f.MODELS
    assert (seg.align_corners, seg.num_classes, seg.out_channels) == (False, 150, 150)   # encoder_decoder.py:104-106
    n_params = sum(p.numel() for p in seg.parameters())
    assert n_params == 34361112                                                           # 34.36 M
    assert len(seg.backbone.state_dict()) == 823


def test_state_dict_names_cover_reference_checkpoint_keys():
    """golden_tiny.pt's calibration keys were written from the REFERENCE's own state_dict names."""
    g = torch.load(os.path.join(ROOT, "tests", "golden", "golden_tiny.pt"))
    seg = s2f.build_segmentor(s2f.configs.tiny())
    sd = seg.state_dict()
    assert set(g["calib"]) <= set(sd)
    seg.load_state_dict(synth.random_state(s2f.configs.tiny()), strict=True)


def test_error_conventions():
    with pytest.raises(AssertionError):                       # sdtv2.py:271-273
        s2f.Spiking_vit_MetaFormer(embed_dim=[64, 128, 250, 360], num_heads=8, mlp_ratios=4, in_channels=3)
    cfg = s2f.configs.tiny()
    cfg["decode_head"]["pixel_decoder"]["encoder"]["layer_cfg"]["self_attn_cfg"]["group"] = 7
    with pytest.raises(ValueError):                           # dcnv3.py:125-127
        s2f.build_segmentor(cfg)
    seg = s2f.build_segmentor(s2f.configs.tiny())
    with pytest.raises(RuntimeError):                         # no CPU path
        seg.encode_decode(torch.zeros(1, 3, 64, 64))
    for m in seg.modules():                                   # ResetModelHook duck-typing (resetmodel_hook.py:17-37)
        if hasattr(m, "reset"):
            m.reset()
    names = dict(seg.named_modules())
    assert "backbone.block3.0.attn.q_spike" in names and "decode_head.pixel_decoder.mask_feature_spike" in names


def test_bn_fold_and_repconv_reparameterisation_match_the_oracle():
    """fold.repconv_dense3x3: conv1x1 -> BN+pad -> dw3x3 -> conv1x1 -> BN -> BN == ONE dense 3x3 (zero pad)."""
    cfg = s2f.configs.tiny()
    P = synth.synthetic_checkpoint("tiny", cfg)
    key = "backbone.block3.0.attn.q_conv"
    c = P[key + ".1.weight"].numel()
    g = torch.Generator().manual_seed(0)
    x = torch.randint(0, 9, (2, c, 9, 7), generator=g).float() / 8
    ref = port.rep_conv(port.Ctx(P), key, x)
    sd = {k[len("backbone."):]: v for k, v in P.items() if k.startswith("backbone.")}
    wm, b = fold.repconv_dense3x3(sd, key[len("backbone."):])
    w = wm.reshape(c, 3, 3, c).permute(0, 3, 1, 2)
    got = F.conv2d(x.double(), w, b, padding=1)
    assert (got - ref.double()).abs().max() < 2e-5 * ref.abs().max()
    # plain conv + BN fold
    s, t = fold.conv_bn(sd, "downsample2.encode_conv", "downsample2.encode_bn")
    xin = torch.randn(1, sd["downsample2.encode_conv.weight"].shape[1], 8, 8, generator=g)
    ref = port.downsample(port.Ctx(P), "backbone.downsample2", xin, 2, 1, True)
    got = F.conv2d(xin.double(), sd["downsample2.encode_conv.weight"].double(), None, 2, 1) * s.view(1, -1, 1, 1) + t.view(1, -1, 1, 1)
    assert (got - ref.double()).abs().max() < 1e-5 * ref.abs().max()


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "s2f.h")).read()
    declared = set(re.findall(r"S2F_API\s+[\w\s\*]+?\b(s2f_\w+)\s*\(", hdr))
    assert len(declared) >= 14
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert os.path.exists(_lib.LIB_PATH), "libs2f.so not built: run __graft_entry__.build()"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    lib = _lib.lib()                       # loads, sets signatures, checks the ABI version; no compute call
    assert lib.s2f_abi_version() == _lib.ABI_VERSION
    assert ctypes.sizeof(_lib.ConvArgs) == 144 or ctypes.sizeof(_lib.ConvArgs) > 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "spike2former_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_fused_stem_host_folding_reproduces_the_oracle_on_cpu():
    """engine.StemU8 (host side of s2f_stem_u8): digit planes of W / std over the raw bytes + the border-tabulated shift.
    The kernel's arithmetic is emulated here with exact integer matmuls; the result must equal reference preprocessing +
    float64 convolution + BatchNorm (oracle) to 2e-5 of scale, for both memory orders, including every border class."""
    import torch.nn.functional as F

    from oracle import port, weights
    from spike2former_b200 import configs, engine, fold

    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    cfg = configs.tiny()
    P = weights.calibrated_state(cfg, 64, 64)
    sd = {k[len("backbone."):]: v for k, v in P.items() if k.startswith("backbone.downsample1_1.")}
    g = torch.Generator().manual_seed(77)
    H, W = 21, 18
    imgs = [torch.randint(0, 256, (3, H, W), generator=g, dtype=torch.uint8) for _ in range(2)]
    x = port.data_preprocess(imgs, mean=mean, std=std, bgr_to_rgb=True).double()
    s_, t_ = fold.conv_bn(sd, "downsample1_1.encode_conv", "downsample1_1.encode_bn")
    ref = F.conv2d(x, sd["downsample1_1.encode_conv.weight"].double(), stride=2, padding=3)
    ref = ref * s_.view(1, -1, 1, 1) + t_.view(1, -1, 1, 1)                       # [n, Cout, Ho, Wo]
    Ho, Wo = ref.shape[-2:]
    for chw in (True, False):
        stem = engine.StemU8(sd, mean, std, True, chw, torch.device("cpu"))
        cout = stem.cout
        dig = stem.packed.view(3, 64, 256)[:, :cout, :192].to(torch.int64)
        wint = dig[0] * 16384 + dig[1] * 128 + dig[2]                             # exact integer weights [Cout, 192]
        tab = stem.tab.view(4, 4, 4, 4, cout).double()
        raw = torch.stack(imgs).to(torch.int64)                                  # [n, 3, H, W] stored (BGR) order
        out = torch.zeros(2, cout, Ho, Wo, dtype=torch.float64)
        for ho in range(Ho):
            for wo in range(Wo):
                a = torch.zeros(2, 192, dtype=torch.int64)
                for kh in range(7):
                    y = 2 * ho - 3 + kh
                    for kw in range(7):
                        xx = 2 * wo - 3 + kw
                        if 0 <= y < H and 0 <= xx < W:
                            for c in range(3):
                                j = c * 7 + kw if chw else kw * 3 + c
                                a[:, kh * 24 + j] = raw[:, c, y, xx]
                top, bot = min(max(3 - 2 * ho, 0), 3), min(max(2 * ho + 4 - H, 0), 3)
                lef, rig = min(max(3 - 2 * wo, 0), 3), min(max(2 * wo + 4 - W, 0), 3)
                S = (a @ wint.t()).double()                                       # the int32 accumulators, exactly
                out[:, :, ho, wo] = S * stem.scale.double()[None, :] + tab[top, bot, lef, rig][None, :]
        scale = max(1.0, ref.abs().max().item())
        assert (out - ref).abs().max().item() < 2e-5 * scale, chw


REF_CFG = "/root/reference/Segmentation/configs/Spike2Former"


@pytest.mark.skipif(not os.path.isdir(REF_CFG), reason="the reference tree only exists in the build container")
@pytest.mark.parametrize("fname,params,classes,T", [
    ("SDTv2_maskformer_DCNpixelDecoder_ade20k.py", 34361112, 150, 1),
    ("SDTv2_maskformer_DCNPixelDecoder_CityScapes.py", 37491605, 19, 1),
    ("SDTv2_maskformer_cocostuff10k_512x512.py", None, 171, 4),
])
def test_reference_config_files_drop_in_unchanged(fname, params, classes, T):
    """The reference's own config files (exec'd as plain Python: `_base_` is just a list literal) -> their `model=`
    dict -> MODELS.build: the registry surface of BASELINE.json's north star ("the existing configs under
    Segmentation/configs drop in unchanged")."""
    ns = {}
    exec(compile(open(os.path.join(REF_CFG, fname)).read(), fname, "exec"), ns)
    seg = s2f.MODELS.build(ns["model"])
    assert type(seg).__name__ == "EncoderDecoder" and type(seg.backbone).__name__ == "Spiking_vit_MetaFormer"
    assert type(seg.decode_head.pixel_decoder).__name__ == "DCNTransformerEncoderPixelDecoder"
    assert seg.num_classes == classes and seg.backbone.T == T
    assert seg.backbone.init_cfg["checkpoint"] == ns["checkpoint_file"]           # kept for init_weights (sdtv2.py:577-612)
    if params is not None:
        assert sum(p.numel() for p in seg.parameters()) == params
    dp = seg.data_preprocessor
    assert dp.channel_conversion and dp.size == tuple(ns["crop_size"]) or list(dp.size) == list(ns["crop_size"])
    if "CityScapes" in fname:
        assert dp.test_cfg == dict(size_divisor=32)
        mine = s2f.configs.cityscapes()["data_preprocessor"]
        assert tuple(mine["size"]) == tuple(ns["crop_size"]) and mine["test_cfg"] == dict(size_divisor=32)


def test_preprocessor_metainfo_follows_stack_batch():
    """SegDataPreProcessor.forward must write the metainfo keys of mmseg's stack_batch
    (data_preprocessor.py:127-150, utils/misc.py:94-116): training -> img_shape (UNPADDED), pad_shape, padding_size;
    test-time padding -> img_padding_size and pad_shape only, img_shape untouched.  CPU part: the bookkeeping is
    checked with the normalisation kernel stubbed out (no GPU here)."""
    from spike2former_b200 import models

    class Field:
        def __init__(self, data):
            self.data = data

        @property
        def shape(self):
            return self.data.shape

    class Sample:
        def __init__(self, h, w):
            self.gt_sem_seg = Field(torch.zeros(1, h, w, dtype=torch.long))
            self.meta = {"img_shape": (h, w), "ori_shape": (h, w)}

        def __contains__(self, k):
            return hasattr(self, k)

        def set_metainfo(self, d):
            self.meta.update(d)

    def fake_normalized(self, batch, size=None, size_divisor=None, out=None):
        hp, wp = self._padded_size(batch.shape[2], batch.shape[3], size, size_divisor)
        return torch.zeros(batch.shape[0], hp, wp, 3)

    pre = models.SegDataPreProcessor(mean=[1, 2, 3], std=[1, 1, 1], size=(64, 96), bgr_to_rgb=True, seg_pad_val=255,
                                     test_cfg=dict(size_divisor=32))
    orig_norm, orig_cuda = models.SegDataPreProcessor.normalized, torch.Tensor.cuda
    models.SegDataPreProcessor.normalized = fake_normalized
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        imgs = [torch.zeros(3, 50, 70, dtype=torch.uint8) for _ in range(2)]
        tr = [Sample(50, 70) for _ in range(2)]
        out = pre({"inputs": imgs, "data_samples": tr}, training=True)
        assert tuple(out["inputs"].shape) == (2, 3, 64, 96)
        for s in tr:
            assert tuple(s.meta["img_shape"]) == (50, 70) and tuple(s.meta["pad_shape"]) == (1, 64, 96)
            assert tuple(s.meta["padding_size"]) == (0, 26, 0, 14) and "img_padding_size" not in s.meta
            assert tuple(s.gt_sem_seg.data.shape) == (1, 64, 96) and int(s.gt_sem_seg.data[0, -1, -1]) == 255
        te = [Sample(50, 70) for _ in range(2)]
        out = pre({"inputs": imgs, "data_samples": te}, training=False)
        assert tuple(out["inputs"].shape) == (2, 3, 64, 96)             # size_divisor 32: 50 x 70 -> 64 x 96
        for s in te:
            assert tuple(s.meta["img_padding_size"]) == (0, 26, 0, 14) and tuple(s.meta["pad_shape"]) == (64, 96)
            assert tuple(s.meta["img_shape"]) == (50, 70) and "padding_size" not in s.meta       # untouched
            assert tuple(s.gt_sem_seg.data.shape) == (1, 50, 70)                                  # labels are not padded at test time
    finally:
        models.SegDataPreProcessor.normalized, torch.Tensor.cuda = orig_norm, orig_cuda
