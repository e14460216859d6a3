import torch, sys
sys.path.insert(0, ".")
import spike2former_b200 as s2f
from spike2former_b200 import synth, engine
cfg = s2f.configs.ade20k()
seg = s2f.build_segmentor(cfg); seg.load_state_dict(synth.synthetic_checkpoint("ade20k", cfg), strict=True); seg = seg.cuda()
g = torch.Generator().manual_seed(11)
H, W = int(sys.argv[1]), int(sys.argv[2])
x = torch.randn(3, 3, H, W, generator=g).cuda()

class Rec(engine.NullProbe):
    active = False
    def __init__(self): self.log = []
recs = {}
def run(xx, tag):
    out = {}
    with torch.no_grad():
        feats = engine.backbone_forward(seg.backbone, xx)
        for i, (s, sp) in enumerate(feats): out[f"bb{i}.s"] = s.clone(); out[f"bb{i}.sp"] = sp.clone()
        mf, memory, ms = engine.pixel_decoder_forward(seg.decode_head.pixel_decoder, feats, want_mask_feature=False)
        out["pd.ysp"] = mf.clone(); out["pd.memory"] = memory.clone()
        for i, m in enumerate(ms): out[f"pd.ms{i}"] = m.clone()
        cls, me, ysp = engine.head_forward(seg.decode_head, feats, last_only=True)
        out["hd.cls"] = cls.clone(); out["hd.me"] = me.clone()
    return out
a = run(x, "b3"); b = run(x[:1].contiguous(), "b1")
for k in a:
    ta, tb = a[k], b[k]
    if k.startswith("hd."):
        ta = ta[:, :1]
    else:
        ta = ta[:1]
    d = (ta.float() - tb.float()).abs()
    print(f"{k:12s} equal={torch.equal(ta, tb)} maxdiff={d.max().item():.4g} nnz={int((d>0).sum())}/{d.numel()}")
