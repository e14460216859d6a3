"""Launches the fused uint8 stem (s2f_stem_u8) three times at batch 32, 512x512, CHW input -- the target of
    ncu --set full -k regex:stem_u8 -s 2 -c 1 -o gpurun_out/stem python tools/prof_stem_u8.py      (GPU box only)"""
import torch, sys
sys.path.insert(0, ".")
import spike2former_b200 as s2f
from spike2former_b200 import engine, ops, synth
cfg = s2f.configs.ade20k(); seg = s2f.build_segmentor(cfg); seg.load_state_dict(synth.synthetic_checkpoint("ade20k", cfg), strict=True); seg = seg.cuda()
x = torch.randint(0, 256, (32, 3, 512, 512), dtype=torch.uint8, device="cuda")
plan = engine.plan_of(seg.backbone, engine.BackbonePlan)
sd = {k: v.detach().cpu() for k, v in seg.backbone.state_dict().items() if k.startswith("downsample1_1.")}
pre = seg.data_preprocessor
st = engine.StemU8(sd, pre._mean_host, pre._std_host, True, True, torch.device("cuda"))
for _ in range(3): st(x)
torch.cuda.synchronize()
