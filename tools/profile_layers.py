"""Per-launch CUDA-event timing of one forward, grouped by (kernel class, layer shape).  GPU box only.
    python tools/profile_layers.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spike2former_b200 as s2f  # noqa: E402
from spike2former_b200 import engine, ops, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ONLY = sys.argv[2] if len(sys.argv) > 2 else ""          # kernel-class filter for the listing
cfg = s2f.configs.ade20k()
seg = s2f.build_segmentor(cfg)
seg.load_state_dict(synth.synthetic_checkpoint("ade20k", cfg), strict=True)
seg = seg.cuda()
x = torch.randint(0, 256, (B, 3, 512, 512), dtype=torch.uint8, device="cuda")          # the uint8 batch bench.py feeds
with torch.no_grad():
    for _ in range(2):
        engine.segmentor_logits(seg, x)
    torch.cuda.synchronize()
    prof = ops.Profiler()
    torch.cuda._sleep(int(6e8))        # GPU parked while the host enqueues: events bracket device time only
    ops.set_profiler(prof)
    engine.segmentor_logits(seg, x)
    ops.set_profiler(None)
torch.cuda.synchronize()
agg = {}
for cls, e0, e1, fl, by, detail, _ex in prof.records:
    a = agg.setdefault((cls, detail), [0.0, 0.0, 0.0, 0])
    a[0] += e0.elapsed_time(e1); a[1] += fl; a[2] += by; a[3] += 1
tot = sum(a[0] for a in agg.values())
print(f"batch {B}: {tot:.2f} ms in {sum(a[3] for a in agg.values())} launches")
rows = [kv for kv in sorted(agg.items(), key=lambda kv: -kv[1][0]) if ONLY in kv[0][0]]
for (cls, detail), a in rows[:60]:
    tf = a[1] / a[0] / 1e9 if a[0] > 0 else 0
    gb = a[2] / a[0] / 1e6 if a[0] > 0 else 0
    print(f"{a[0]:8.3f} ms {100 * a[0] / tot:5.1f}%  x{a[3]:<3d} {cls:16s} {detail:58s} {tf:8.1f} TFLOP/s {gb:8.0f} GB/s")
