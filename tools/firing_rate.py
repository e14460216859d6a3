"""Per-neuron firing-rate / mean-level census of one forward (GPU box only) -- the B200 counterpart of the reference's
tools/cal_firing_num.py:140-171, which hooks every neuron and averages its fp32 output.  Here the int8 levels the
kernels already emit are histogrammed on the device (s2f_level_hist), 1 byte read per neuron.

    python tools/firing_rate.py [ade20k|cityscapes|tiny] [batch] [HxW]
Prints one line per neuron and the totals the paper's energy model needs (spikes = sum of levels, i.e. the number of
unit spikes an integer-valued neuron stands for, and synaptic-operation counts are spikes x fan-out)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spike2former_b200 as s2f  # noqa: E402
from spike2former_b200 import engine, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ade20k"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
H, W = (int(v) for v in sys.argv[3].split("x")) if len(sys.argv) > 3 else ((64, 64) if name == "tiny" else (512, 512))
cfg = getattr(s2f.configs, name)()
seg = s2f.build_segmentor(cfg)
seg.load_state_dict(synth.synthetic_checkpoint(name if name != "tiny" else "ade20k", cfg) if name != "tiny"
                    else __import__("oracle.weights", fromlist=["x"]).calibrated_state(cfg, H, W), strict=True)
seg = seg.cuda()
g = torch.Generator().manual_seed(0)
img = torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8).cuda()
census = engine.FiringCensus()
with torch.no_grad():
    engine.segmentor_logits(seg, img, census)
rep = census.report()
tot_n = sum(r["elements"] for r in rep.values())
tot_spikes = sum(sum(i * c for i, c in enumerate(r["hist"])) for r in rep.values())
for k, r in rep.items():
    print(f"{k:70s} n={r['elements']:>11d} rate={r['firing_rate']:.4f} mean_level={r['mean_level']:.4f} hist={r['hist'][:9]}")
print(json.dumps(dict(config=name, batch=B, size=[H, W], neurons=len(rep), elements=tot_n, unit_spikes=tot_spikes,
                      mean_firing_rate=sum(r["firing_rate"] * r["elements"] for r in rep.values()) / tot_n,
                      mean_level=tot_spikes / tot_n)))
