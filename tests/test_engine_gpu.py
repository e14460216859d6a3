"""GPU parity: the CUDA engine, driven unit by unit with the oracle's tensors (teacher forcing)."""
import pytest
import torch

import spike2former_b200 as s2f
from oracle import port, probe, weights

pytestmark = pytest.mark.gpu


def _run(cfg, H, W, force=True, batch=1):
    torch.manual_seed(0)
    P = weights.calibrated_state(cfg, H, W)
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
