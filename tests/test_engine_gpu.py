"""GPU parity: the CUDA engine, driven unit by unit with the oracle's tensors (teacher forcing)."""
import pytest
import torch

import spike2former_b200 as s2f
from oracle import port, probe, weights

pytestmark = pytest.mark.gpu


def _run(cfg, H, W, force=True, batch=1, style="default"):
    torch.manual_seed(0)
    P = weights.calibrated_state(cfg, H, W, style=style)
    img = weights.test_image(cfg, H, W, batch)
    taps, marks, ref_logits = probe.record_oracle(P, cfg, img)
    seg = s2f.build_segmentor(cfg)
    seg.load_state_dict(P, strict=True)
    seg = seg.cuda()
    pr = probe.TeacherProbe(taps, marks, torch.device("cuda"), force=force)
    from spike2former_b200 import engine

    with torch.no_grad():
        logits = engine.segmentor_logits(seg, img.cuda(), pr)
    torch.cuda.synchronize()
    return pr, logits.cpu(), ref_logits, taps


def _report(pr):
    s = pr.summary()
    worst = sorted((e for e in pr.log if e["kind"] == "spike" and "flips" in e), key=lambda e: -e["flips"])[:5]
    print("summary", s)
    print("  worst flip gap (relative distance of a flipped pre-activation from k+0.5):",
          max([e.get("worst_gap", 0.0) for e in pr.log if e["kind"] == "spike"] or [0.0]))
    for e in pr.log:
        if e["kind"] == "spike" and e.get("unexplained", 0) > 0:
            print("  UNEXPLAINED:", e)
    for e in worst:
        print("  most flips:", e)
    for e in sorted((e for e in pr.log if e["kind"] == "real" and "rel" in e), key=lambda e: -e["rel"])[:5]:
        print("  worst real:", e)
    return s


def test_tiny_teacher_forced():
    cfg = s2f.configs.tiny()
    pr, logits, ref, taps = _run(cfg, 64, 64)
    s = _report(pr)
    assert s["unknown"] == [], s["unknown"]
    assert s["neurons"] >= len(taps) - 1            # encoder_in_proj_spike is never called by the reference
    assert s["unexplained"] == 0
    assert s["maxdev"] <= 1
    assert s["flips"] <= 1e-4 * s["spike_elems"]
    assert s["worst_rel"] < 1e-4, s["worst_real"]
    rel = float((logits - ref).abs().max() / ref.abs().max())
    print("teacher-forced logits rel err", rel)
    assert rel < 1e-2
    agree = float((logits.argmax(1) == ref.argmax(1)).float().mean())
    print("argmax agreement", agree, "classes", ref.argmax(1).unique().numel())
    assert agree >= 0.999


def test_tiny_T2_teacher_forced():
    """T > 1 (the T = 4 cocostuff configs of the reference): T folds into the batch, the head averages over it."""
    cfg = s2f.configs.tiny()
    cfg["backbone"]["T"] = 2
    pr, logits, ref, taps = _run(cfg, 64, 64, batch=2)
    s = _report(pr)
    assert s["unknown"] == [] and s["unexplained"] == 0 and s["maxdev"] <= 1
    assert logits.shape == ref.shape == (2, 11, 64, 64)
    assert float((logits - ref).abs().max() / ref.abs().max()) < 1e-2
    assert float((logits.argmax(1) == ref.argmax(1)).float().mean()) >= 0.999


def test_ade20k_teacher_forced_512():
    """The graded configuration (SDTv2 + DCN pixel decoder, ADE20K shape 512x512), unit-by-unit parity."""
    cfg = s2f.configs.ade20k()
    pr, logits, ref, taps = _run(cfg, 512, 512)
    s = _report(pr)
    assert s["unknown"] == []
    assert s["neurons"] == 270 and s["spike_elems"] == 187913216       # SURVEY.md appendix A census
    assert s["unexplained"] == 0 and s["maxdev"] <= 1
    assert s["flips"] <= 1e-5 * s["spike_elems"]
    assert s["worst_rel"] < 1e-4, s["worst_real"]
    rel = float((logits - ref).abs().max() / ref.abs().max())
    agree = float((logits.argmax(1) == ref.argmax(1)).float().mean())
    ncls = ref.argmax(1).unique().numel()
    print(f"ADE20K 512: logits rel err {rel:.3e}, argmax agreement {agree:.6f}, classes in oracle argmax {ncls}")
    assert rel < 1e-2 and agree >= 0.999 and ncls >= 20


def test_cityscapes_config_teacher_forced_256x512():
    """Cityscapes config (19 classes, encoder FFN 2048) at a non-square shape: ragged class tile in the tail,
    different GEMM widths, rectangular maps everywhere."""
    cfg = s2f.configs.cityscapes()
    pr, logits, ref, taps = _run(cfg, 256, 512)
    s = _report(pr)
    assert s["unknown"] == [] and s["neurons"] == 270
    assert s["unexplained"] == 0 and s["maxdev"] <= 1
    assert s["flips"] <= 1e-5 * s["spike_elems"]
    assert s["worst_rel"] < 1e-4, s["worst_real"]
    rel = float((logits - ref).abs().max() / ref.abs().max())
    agree = float((logits.argmax(1) == ref.argmax(1)).float().mean())
    print(f"Cityscapes cfg 256x512: logits rel err {rel:.3e}, argmax agreement {agree:.6f}")
    assert rel < 1e-2 and agree >= 0.999


def test_cityscapes_config_teacher_forced_full_1024x2048():
    """BASELINE config 4 at its full shape (batch 1): 1.53 G spike elements, 16384/65536/262144-key cross-attention
    (the 64-bit accumulation path of the attention kernel), 8192-token SDSA, 512 x 1024 mask maps."""
    from spike2former_b200 import engine, synth

    cfg = s2f.configs.cityscapes()
    P = synth.synthetic_checkpoint("cityscapes", cfg)            # calibrated at 1024 x 2048 (tools/make_calibration.py)
    img = weights.test_image(cfg, 1024, 2048)
    taps, marks, ref = probe.record_oracle(P, cfg, img)
    seg = s2f.build_segmentor(cfg)
    seg.load_state_dict(P, strict=True)
    seg = seg.cuda()
    pr = probe.TeacherProbe(taps, marks, torch.device("cuda"))
    with torch.no_grad():
        logits = engine.segmentor_logits(seg, img.cuda(), pr).cpu()
    s = _report(pr)
    assert s["unknown"] == [] and s["neurons"] == 270 and s["spike_elems"] == 1526936576
    assert s["unexplained"] == 0 and s["maxdev"] <= 1
    assert s["flips"] <= 1e-5 * s["spike_elems"]
    assert s["worst_rel"] < 1e-4, s["worst_real"]
    rel = float((logits - ref).abs().max() / ref.abs().max())
    agree = float((logits.argmax(1) == ref.argmax(1)).float().mean())
    ncls = ref.argmax(1).unique().numel()
    print(f"Cityscapes 1024x2048: logits rel err {rel:.3e}, argmax agreement {agree:.6f}, classes {ncls}")
    assert rel < 1e-2 and agree >= 0.999 and ncls >= 10


def test_free_running_report():
    """No forcing: reports how spike flips grow through the (chaotic, random-init) network."""
    cfg = s2f.configs.tiny()
    pr, logits, ref, taps = _run(cfg, 64, 64, force=False)
    sp = [e for e in pr.log if e["kind"] == "spike" and "flips" in e]
    first = sp[:12]
    print("free-running flips per neuron (first 12):", [(e["name"].split(".")[-1], e["flips"]) for e in first])
    print("free-running total flips", sum(e["flips"] for e in sp), "of", sum(e["numel"] for e in sp))
    print("free-running logits rel err", float((logits - ref).abs().max() / ref.abs().max()))
    assert torch.isfinite(logits).all()
    assert sum(e["flips"] for e in sp[:4]) <= 4          # the first units still see identical inputs


def test_cuda_graph_replay_matches_eager_launches():
    """EncoderDecoder.encode_decode / predict_labels replay a captured CUDA graph: same bits as launch-by-launch."""
    from spike2former_b200 import engine, synth

    cfg = s2f.configs.tiny()
    seg = s2f.build_segmentor(cfg)
    seg.load_state_dict(synth.synthetic_checkpoint("tiny", cfg), strict=True)
    seg = seg.cuda()
    seg.graph_min_hits = 1
    g = torch.Generator().manual_seed(3)
    x1, x2 = (torch.randn(2, 3, 64, 64, generator=g).cuda() for _ in range(2))
    with torch.no_grad():
        e1 = engine.segmentor_logits(seg, x1).clone()
        e2 = engine.segmentor_logits(seg, x2).clone()
        g1 = seg.encode_decode(x1).clone()
        g2 = seg.encode_decode(x2).clone()
        g1b = seg.encode_decode(x1).clone()
        lab = seg.predict_labels(x2).clone()
    assert len(seg._graphs) == 2 and all(gr.launches > 50 for gr in seg._graphs.values())
    assert torch.equal(g1, e1) and torch.equal(g2, e2) and torch.equal(g1b, e1)
    assert not torch.equal(e1, e2)
    assert torch.equal(lab.long(), e2.argmax(1))
    seg.use_cuda_graph = False
    with torch.no_grad():
        assert torch.equal(seg.encode_decode(x1), e1)


@pytest.mark.parametrize("B,H,W", [(1, 512, 512), (3, 384, 640), (2, 256, 128)])
def test_public_api_shapes_and_batches(B, H, W):
    """Public path (EncoderDecoder.encode_decode / predict_labels, CUDA-graph replay) at other batches and
    rectangular shapes of the ADE20K model: finite logits, labels == argmax(logits), images independent of the batch."""
    from spike2former_b200 import synth

    cfg = s2f.configs.ade20k()
    seg = s2f.build_segmentor(cfg)
    seg.load_state_dict(synth.synthetic_checkpoint("ade20k", cfg), strict=True)
    seg = seg.cuda()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 3, H, W, generator=g).cuda()
    with torch.no_grad():
        logits = seg.encode_decode(x).clone()
        labels = seg.predict_labels(x).clone()
        single = seg.encode_decode(x[:1].contiguous()).clone()
    assert logits.shape == (B, 150, H, W) and labels.shape == (B, H, W)
    assert torch.isfinite(logits).all()
    assert torch.equal(labels.long(), logits.argmax(1))
    assert torch.equal(single[0], logits[0])                 # batch-sharding invariant: an image does not see its batch
    assert logits.argmax(1).unique().numel() >= 2                 # not a constant map (small random-weight images show few classes)


def test_graph_cache_follows_weight_changes_and_is_bounded():
    """a captured graph holds raw pointers into the plans; any parameter change -- through a CHILD's
    load_state_dict or an in-place edit -- must re-capture, never replay stale memory; one-off shapes run eagerly;
    the cache is an LRU; results are fresh tensors unless aliasing is requested."""
    from spike2former_b200 import engine, synth

    cfg = s2f.configs.tiny()
    seg = s2f.build_segmentor(cfg)
    P = synth.synthetic_checkpoint("tiny", cfg)
    seg.load_state_dict(P, strict=True)
    seg = seg.cuda()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 3, 64, 64, generator=g).cuda()
    with torch.no_grad():
        a0 = seg.encode_decode(x)                      # first sighting of the shape: eager
        assert len(seg._graphs) == 0
        a1 = seg.encode_decode(x)                      # second: captured + replayed
        assert len(seg._graphs) == 1 and torch.equal(a0, a1)
        a2 = seg.encode_decode(x)
        assert a2.data_ptr() != a1.data_ptr() and torch.equal(a1, a2)          # fresh result tensors
        # (1) child load_state_dict with different weights
        P2 = synth.random_state(cfg, seed=99)
        for k, v in P.items():
            if "running_" in k:
                P2[k] = v
        seg.backbone.load_state_dict({k[9:]: v for k, v in P2.items() if k.startswith("backbone.")})
        b = seg.encode_decode(x)
        want = engine.segmentor_logits(seg, x)
        assert torch.equal(b, want) and not torch.equal(b, a1)
        # (2) in-place edit of one parameter (class 0 becomes every query's favourite)
        seg.decode_head.cls_embed.bias[0] += 5.0          # (no_grad) bumps the tensor's version counter
        c = seg.encode_decode(x)
        assert torch.equal(c, engine.segmentor_logits(seg, x)) and not torch.equal(c, b)
        # (3) LRU bound
        seg.graph_cache_size, seg.graph_min_hits = 2, 1
        for hw in (32, 64, 96, 128):
            seg.encode_decode(torch.randn(1, 3, hw, hw, generator=g).cuda())
        assert len(seg._graphs) == 2
        # (4) aliasing on request
        seg.alias_graph_output = True
        y = torch.randn(1, 3, 128, 128, generator=g).cuda()
        assert seg.encode_decode(y).data_ptr() == seg.encode_decode(y).data_ptr()
