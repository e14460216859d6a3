"""Generates tests/golden/*.pt by running THE REFERENCE ITSELF (oracle/ref_loader.py executes the files under
/root/reference) in the build container.  Usage:  python tests/golden/make_golden.py

golden_tiny.pt   : tiny config (same topology), 64x64, batch 1: calibration statistics measured by the
                   REFERENCE in train mode, every neuron's level tensor (int8) and the seg logits.
golden_ade20k.pt : the graded ADE20K config at 512x512: per-neuron checksums (sum of levels, number of
                   non-zeros, exact .5 ties) for all 270 neurons, a 64x64 crop of the logits, the argmax histogram.
golden_nilif.pt  : Q_IFNode(Quant()) known answers: tie vectors, clamp edges, stateful second call.
"""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import port, ref_loader, weights  # noqa: E402
from spike2former_b200 import configs, synth  # noqa: E402


def reference_calibrated(cfg, h, w):
    """Calibrate with the REFERENCE modules (train mode, momentum 1), then the port's centring step."""
    P = synth.random_state(cfg)
    bb, hd = ref_loader.build_reference(cfg)
    bb.load_state_dict({k[9:]: v for k, v in P.items() if k.startswith("backbone.")}, strict=True)
    hd.load_state_dict({k[12:]: v for k, v in P.items() if k.startswith("decode_head.")}, strict=True)
    for m in list(bb.modules()) + list(hd.modules()):
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
            m.momentum = 1.0
    bb.train(); hd.train()
    with torch.no_grad():
        ref_loader.reference_predict(bb, hd, weights.calibration_batch(cfg, h, w))
    bb.eval(); hd.eval()
    ref_sd = {**{"backbone." + k: v for k, v in bb.state_dict().items()},
              **{"decode_head." + k: v for k, v in hd.state_dict().items()}}
    P2 = weights.calibrated_state(cfg, h, w)          # port-side calibration (incl. centring)
    for k, v in ref_sd.items():
        if "running_" in k:
            assert torch.equal(v, P2[k]), f"port calibration differs from the reference at {k}"
    bb.load_state_dict({k[9:]: v for k, v in P2.items() if k.startswith("backbone.")}, strict=True)
    hd.load_state_dict({k[12:]: v for k, v in P2.items() if k.startswith("decode_head.")}, strict=True)
    return P2, bb, hd


def run_reference(bb, hd, img):
    tap = ref_loader.SpikeTap(bb, hd)
    with torch.no_grad():
        logits = ref_loader.reference_predict(bb, hd, img)
    tap.close()
    return logits, tap.records


def main():
    torch.set_num_threads(8)
    # ---- tiny
    cfg = configs.tiny()
    P, bb, hd = reference_calibrated(cfg, 64, 64)
    img = weights.test_image(cfg, 64, 64)
    logits, recs = run_reference(bb, hd, img)
    torch.save(dict(calib=synth.calibration_of(P), mask_gain=weights.MASK_GAIN, logits=logits,
                    levels={n: (o * 8).round().to(torch.int8) for n, x, o in recs},
                    ties={n: port.count_ties(x) for n, x, o in recs}), os.path.join(HERE, "golden_tiny.pt"))
    print("tiny:", len(recs), "neurons, logits", tuple(logits.shape))
    # ---- ade20k 512
    cfg = configs.ade20k()
    P, bb, hd = reference_calibrated(cfg, 512, 512)
    shipped = torch.load(synth.calibration_path("ade20k"))["calib"]
    for k, v in synth.calibration_of(P).items():
        assert torch.equal(v, shipped[k]), f"shipped calibration differs at {k}"
    img = weights.test_image(cfg, 512, 512)
    logits, recs = run_reference(bb, hd, img)
    am = logits.argmax(1)
    torch.save(dict(checks={n: (int((o * 8).round().sum()), int((o > 0).sum()), port.count_ties(x), tuple(x.shape))
                            for n, x, o in recs},
                    logits_crop=logits[:, :, 224:288, 224:288].clone(), argmax_hist=torch.bincount(am.flatten(), minlength=150),
                    logits_absmax=float(logits.abs().max())), os.path.join(HERE, "golden_ade20k.pt"))
    print("ade20k:", len(recs), "neurons,", sum(x.numel() for _, x, _ in recs), "spike elements,",
          am.unique().numel(), "classes in argmax")
    # ---- neuron known answers from the reference's own Q_IFNode
    ns = ref_loader.load()
    node = ns.neuron.Q_IFNode(surrogate_function=ns.surrogate.Quant())
    g = torch.Generator().manual_seed(5)
    x = torch.cat([torch.tensor([k + 0.5 for k in range(9)] + [-0.0, 0.0, -1.0, 8.0, 8.25, 9.0, 100.0, -100.0, 7.999999]),
                   torch.rand(4096, generator=g) * 12 - 2])
    node.reset(); y1 = node(x).clone(); y2 = node(x).clone(); v2 = node.v.clone()     # second call carries the membrane
    torch.save(dict(x=x, y1=y1, y2=y2, v2=v2), os.path.join(HERE, "golden_nilif.pt"))
    print("nilif:", x.numel(), "values; first five levels", (y1[:5] * 8).tolist())


if __name__ == "__main__":
    main()
