"""CPU: the training-step oracle (oracle/port.py in train mode + oracle/train_port.py) against THE REFERENCE's own
training forward, losses, Hungarian matching and autograd (tests/golden/make_golden_train.py -> golden_train.pt)."""
import importlib.util
import os

import pytest
import torch

from oracle import ref_loader, train_port, weights
from spike2former_b200 import configs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _maker():
    spec = importlib.util.spec_from_file_location("make_golden_train", os.path.join(GOLD, "make_golden_train.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _port_run(cfg, P, img, gt):
    P = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v.clone()) for k, v in P.items()}
    losses = train_port.train_losses(P, cfg, img, gt)
    train_port.total_loss(losses).backward()
    return losses, P


def _compare(losses, P, gold, tol):
    for k, v in gold["losses"].items():
        assert abs(float(losses[k]) - float(v)) <= tol * max(1.0, abs(float(v))), (k, float(losses[k]), float(v))
    worst = 0.0
    for k, n in gold["grad_norm"].items():
        g = P[k].grad
        assert g is not None, k
        scale = max(float(n), 1e-6)
        worst = max(worst, abs(float(g.norm()) - float(n)) / scale,
                    float((g.reshape(-1)[:8] - gold["grad_head"][k]).abs().max()) / scale)
    for k, v in gold["stats"].items():
        assert torch.allclose(P[k], v, rtol=1e-5, atol=1e-7), k        # running statistics were updated with momentum 0.1
    return worst


def test_training_oracle_matches_reference_golden():
    """Losses (21 terms) and the gradient of all 806 parameters, from the reference-generated fixture."""
    gold = torch.load(os.path.join(GOLD, "golden_train.pt"))
    cfg = configs.tiny()
    img, gt = _maker().inputs(cfg)
    losses, P = _port_run(cfg, weights.calibrated_state(cfg, 64, 64, style="stable"), img, gt)
    assert set(losses) == set(gold["losses"]) and len(gold["grad_norm"]) == 806
    # identical on the build container's CPU; another host ISA may reorder sums (a near-tie spike flip moves a
    # gradient by more than rounding, so the bound is loose there)
    worst = _compare(losses, P, gold, 1e-4)
    print("worst relative gradient deviation from the reference fixture:", worst)
    assert worst <= 1e-3


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference (build container only)")
def test_training_oracle_against_live_reference_other_seed():
    cfg = configs.tiny()
    mk = _maker()
    img, gt = mk.inputs(cfg, seed=11)
    P0 = weights.calibrated_state(cfg, 64, 64, style="stable")
    ref_losses, ref_grads, ref_stats = mk.reference_run(cfg, P0, img, gt)
    losses, P = _port_run(cfg, P0, img, gt)
    for k, v in ref_losses.items():
        assert abs(float(losses[k]) - float(v)) <= 1e-5 * max(1.0, abs(float(v))), k
    for k, g in ref_grads.items():
        assert (P[k].grad - g).abs().max() <= 1e-4 * max(float(g.abs().max()), 1e-6), k
    for k, v in ref_stats.items():
        assert torch.allclose(P[k], v, rtol=1e-5, atol=1e-7), k
