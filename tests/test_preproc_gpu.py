"""GPU parity of the callers either side of the path: uint8 preprocessing and the
spike-level census, through the C ABI, against the oracle and the reference-generated fixture."""
import os

import pytest
import torch

import spike2former_b200 as s2f
from oracle import port, weights
from spike2former_b200 import engine, ops, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]


def test_preprocess_u8_matches_reference_golden_bitwise():
    g = torch.load(os.path.join(GOLD, "golden_preproc.pt"))
    for c in g["cases"]:
        kw = dict(c["kw"])
        tc = kw.pop("test_cfg", None) or {}
        kw.pop("size", None)
        pp = s2f.models.SegDataPreProcessor(test_cfg=tc or None, **kw)
        out = pp(dict(inputs=[i.cuda() for i in c["imgs"]]), training=False)["inputs"]
        assert out.shape == c["out"].shape
        assert torch.equal(out.cpu(), c["out"]), c["kw"]
        # HWC input (decoder order) gives the same batch
        hwc = c["imgs"].permute(0, 2, 3, 1).contiguous().cuda()
        assert torch.equal(pp(dict(inputs=hwc))["inputs"].cpu(), c["out"])
    out = ops.preprocess_u8(g["ramp"][None].cuda(), mean=MEAN, std=STD, swap_rb=True)
    assert torch.equal(out.permute(0, 3, 1, 2).cpu(), g["table"])            # all 3 x 256 byte values


@pytest.mark.parametrize("H,W,size", [(512, 512, None), (37, 53, (40, 64)), (1, 1, None), (9, 2, (9, 5))])
def test_preprocess_u8_vs_oracle(H, W, size):
    gen = torch.Generator().manual_seed(5)
    imgs = [torch.randint(0, 256, (3, H, W), generator=gen, dtype=torch.uint8) for _ in range(3)]
    want = port.data_preprocess(imgs, mean=MEAN, std=STD, bgr_to_rgb=True, size=size, pad_val=0.25)
    got = ops.preprocess_u8(torch.stack(imgs).cuda(), mean=MEAN, std=STD, swap_rb=True, size=size, pad_val=0.25)
    assert torch.equal(got.permute(0, 3, 1, 2).cpu(), want)


def test_preprocess_rejects_cpu_and_wrong_dtype():
    with pytest.raises(RuntimeError):
        ops.preprocess_u8(torch.zeros(1, 3, 4, 4, dtype=torch.uint8))
    with pytest.raises(RuntimeError):
        ops.preprocess_u8(torch.zeros(1, 3, 4, 4, device="cuda"))


@pytest.mark.parametrize("N", [0, 5, 16, 1000, 1 << 20, (1 << 20) + 7])
def test_level_hist_exact(N):
    gen = torch.Generator().manual_seed(N)
    lv = torch.randint(0, 9, (N,), generator=gen, dtype=torch.int8)
    hist = ops.level_hist(lv.cuda())
    want = torch.bincount(lv.long(), minlength=16)
    assert torch.equal(hist.cpu(), want)
    hist = ops.level_hist(lv.cuda(), hist)                       # accumulates
    assert torch.equal(hist.cpu(), 2 * want)


def _tiny_segmentor():
    cfg = s2f.configs.tiny()
    seg = s2f.build_segmentor(cfg)
    seg.load_state_dict(weights.calibrated_state(cfg, 64, 64), strict=True)
    seg.graph_min_hits = 1            # capture at the first sighting of a shape: these tests are about graph replay
    return seg.cuda()


def test_uint8_images_through_the_segmentor_equal_preprocessed_fp32(monkeypatch):
    """Separate preprocessing kernel: encode_decode(uint8 batch) == encode_decode(reference-preprocessed fp32 batch),
    bit for bit, eager and graphed."""
    monkeypatch.setattr(engine, "FUSE_STEM_U8", False)
    seg = _tiny_segmentor()
    gen = torch.Generator().manual_seed(8)
    imgs = [torch.randint(0, 256, (3, 64, 64), generator=gen, dtype=torch.uint8) for _ in range(2)]
    x = port.data_preprocess(imgs, mean=MEAN, std=STD, bgr_to_rgb=True).cuda()
    u8 = torch.stack(imgs).cuda()
    with torch.no_grad():
        want = engine.segmentor_logits(seg, x).clone()
        got = engine.segmentor_logits(seg, u8).clone()
        assert torch.equal(got, want)
        assert torch.equal(seg.encode_decode(u8), want)          # CUDA-graph replay with a uint8 static input
        assert torch.equal(seg.predict_labels(u8).long(), want.argmax(1))


@pytest.mark.parametrize("H,W,layout", [(64, 64, "chw"), (64, 64, "hwc"), (37, 53, "chw"), (512, 512, "hwc"), (8, 24, "chw")])
def test_fused_uint8_stem_vs_oracle_and_fp32_stem(H, W, layout):
    """s2f_stem_u8 (preprocessor folded into an int8 tensor-core stem) against (i) the oracle: reference preprocessing
    + float64 convolution + BN, (ii) the fp32 stem kernel fed by the preprocessing kernel.  The pre-activations agree
    to 2e-5 of scale; levels may differ only where the pre-activation is within 1e-4 of a rounding boundary."""
    import torch.nn.functional as F

    from spike2former_b200 import fold

    seg = _tiny_segmentor()
    bb = seg.backbone
    gen = torch.Generator().manual_seed(H * 1000 + W)
    imgs = [torch.randint(0, 256, (3, H, W), generator=gen, dtype=torch.uint8) for _ in range(2)]
    imgs[0][:, :4, :4] = 255
    imgs[1][:, -4:, -4:] = 0
    u8 = torch.stack(imgs)
    dev_in = (u8 if layout == "chw" else u8.permute(0, 2, 3, 1).contiguous()).cuda()
    sd = {k: v.detach().cpu() for k, v in bb.state_dict().items()}
    x = port.data_preprocess(imgs, mean=MEAN, std=STD, bgr_to_rgb=True).double()
    s_, t_ = fold.conv_bn(sd, "downsample1_1.encode_conv", "downsample1_1.encode_bn")
    ref = F.conv2d(x, sd["downsample1_1.encode_conv.weight"].double(), stride=2, padding=3)
    ref = (ref * s_.view(1, -1, 1, 1) + t_.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
    stem = engine.StemU8(sd, MEAN, STD, True, layout == "chw", torch.device("cuda"))
    of, os_ = stem(dev_in)
    of, os_ = of.cpu(), os_.cpu()
    assert of.shape == ref.shape
    scale = max(1.0, ref.abs().max().item())
    assert (of.double() - ref).abs().max().item() < 2e-5 * scale
    assert torch.equal(os_, torch.round(torch.clamp(of, 0, 8)).to(torch.int8))
    want_lv = torch.round(torch.clamp(ref, 0, 8)).to(torch.int8)
    flips = os_ != want_lv
    assert int((flips & ((ref - ref.floor() - 0.5).abs() > 1e-4)).sum()) == 0
    # (ii) the unfused path of this library
    pre = seg.data_preprocessor
    xf = pre.normalized(dev_in)
    plan = engine.plan_of(bb, engine.BackbonePlan)
    f2, s2 = plan.layers["stem"](xf, 2, H, W, f32=True, spike=True)
    assert (of - f2.cpu()).abs().max().item() < 2e-5 * scale
    d = os_ != s2.cpu()
    assert int((d & ((ref - ref.floor() - 0.5).abs() > 1e-4)).sum()) == 0


def test_fused_uint8_stem_end_to_end_graph_equals_eager():
    seg = _tiny_segmentor()
    gen = torch.Generator().manual_seed(9)
    u8 = torch.randint(0, 256, (2, 3, 64, 64), generator=gen, dtype=torch.uint8).cuda()
    with torch.no_grad():
        eager = engine.segmentor_logits(seg, u8).clone()
        assert torch.isfinite(eager).all()
        assert torch.equal(seg.encode_decode(u8), eager)
        assert torch.equal(seg.predict_labels(u8).long(), eager.argmax(1))
        hwc = u8.permute(0, 2, 3, 1).contiguous()
        assert torch.equal(engine.segmentor_logits(seg, hwc), eager)     # same pixels, other memory order


def test_firing_rate_census():
    """engine.FiringCensus (s2f_level_hist over every neuron of one forward): exact counts, and the first neurons'
    firing rates equal the oracle's (later ones drift with the chaotic free-running network)."""
    from oracle import probe

    cfg = s2f.configs.tiny()
    P = weights.calibrated_state(cfg, 64, 64)
    img = weights.test_image(cfg, 64, 64)
    taps, marks, _ = probe.record_oracle(P, cfg, img)
    seg = s2f.build_segmentor(cfg)
    seg.load_state_dict(P, strict=True)
    seg = seg.cuda()
    census = engine.FiringCensus(keep=True)
    with torch.no_grad():
        engine.segmentor_logits(seg, img.cuda(), census)
    rep = census.report()
    assert len(rep) >= 200
    for name, row in rep.items():
        lv = census.kept[name].flatten().long().cpu()
        assert row["elements"] == lv.numel()
        assert row["hist"] == torch.bincount(lv, minlength=16).tolist(), name
    first = "backbone.ConvBlock1_1.0.Conv.spike1"
    pre, lv = taps[first]
    assert abs(rep[first]["firing_rate"] - float((lv != 0).float().mean())) < 1e-3
    assert abs(rep[first]["mean_level"] - 8 * float(lv.float().mean()) * (1 if lv.dtype != torch.int8 else 0.125)) < 1e-2 or \
        abs(rep[first]["mean_level"] - float(lv.float().mean())) < 1e-2
