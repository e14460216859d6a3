"""Build-container tool: measure the calibration statistics of the synthetic checkpoints with the CPU oracle
and write them to spike2former_b200/data/calib_<name>.pt (shipped; loading them needs no oracle code).

    python tools/make_calibration.py ade20k 512 512
    python tools/make_calibration.py ade20k 512 512 stable      # -> calib_ade20k_stable.pt (synth.stable_state)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import weights  # noqa: E402
from spike2former_b200 import configs, synth  # noqa: E402


def main(name, h, w, style="default", seed=1234):
    cfg = getattr(configs, name)()
    P = weights.calibrated_state(cfg, h, w, seed, style=style)
    calib = synth.calibration_of(P)
    calib[synth.CALIBRATED_EXTRA[2]] = torch.tensor(weights.MASK_GAIN_OF[style])
    os.makedirs(synth.DATA_DIR, exist_ok=True)
    name = name if style == "default" else f"{name}_{style}"
    path = synth.calibration_path(name)
    torch.save(dict(seed=seed, h=h, w=w, style=style, calib=calib), path)
    # round trip: seeded weights + file == the calibrated state
    Q = synth.synthetic_checkpoint(name, cfg, seed)
    bad = [k for k in P if not torch.equal(P[k], Q[k])]
    assert not bad, bad[:5]
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB, {len(calib)} tensors; round trip exact")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), *(sys.argv[4:5]))
