"""GPU tool: where does the training step's time go?  torch profiler over one step, top CUDA kernels by total time.
    python tools/profile_train.py [batch] [precision]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spike2former_b200 as s2f  # noqa: E402
from spike2former_b200 import synth, train  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32"
cfg = s2f.configs.ade20k()
seg = s2f.build_segmentor(cfg)
seg.load_state_dict(synth.synthetic_checkpoint("ade20k", cfg), strict=True)
seg = seg.cuda()
step = train.TrainStep(seg, precision=prec)
g = torch.Generator().manual_seed(0)
img = torch.randn(B, 3, 512, 512, generator=g).cuda()
gt = torch.randint(0, 150, (B, 1, 32, 32), generator=g).repeat_interleave(16, 2).repeat_interleave(16, 3).cuda()
for _ in range(2):
    step(img, gt)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    step(img, gt)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
