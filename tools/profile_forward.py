"""GPU tool for ncu: one forward of the benchmarked path inside a cudaProfilerStart/Stop bracket.

    ncu --profile-from-start off ... python tools/profile_forward.py [ade20k|cityscapes] [batch] [logits|labels] [micro|top16]

The model takes the same uint8 batch bench.py feeds it; `micro` profiles the BASELINE config 2 kernels instead
(NI-LIF D=8 / D=4 / folded BN, spike GEMM fc1 512->2048, SDSA C=512 d=64)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spike2former_b200 as s2f  # noqa: E402
from spike2former_b200 import engine, ops, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ade20k"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
labels = (sys.argv[3] if len(sys.argv) > 3 else "logits") == "labels"
micro = len(sys.argv) > 4 and sys.argv[4] == "micro"
g = torch.Generator().manual_seed(0)
if micro:
    x = (torch.rand(64, 1024, 512, generator=g) * 12 - 2).cuda()
    lv = torch.empty(x.shape, dtype=torch.int8, device="cuda")
    sc, sh = (torch.rand(512, generator=g) + 0.5).cuda(), torch.randn(512, generator=g).cuda()
    a = torch.randint(0, 9, (64, 32, 32, 512), generator=g, dtype=torch.int8).cuda()
    w = torch.randn(2048, 512, generator=g) / 512 ** 0.5
    packed, rowscale = ops.pack_weights_i8(w, 1, 512, 3)
    packed, gsc, gsh = packed.cuda(), (rowscale / 8).cuda(), torch.zeros(2048).cuda()
    q, k, v = (torch.randint(0, 3, (64, 1024, 512), generator=g, dtype=torch.int8).cuda() for _ in range(3))

    def run():
        ops.nilif(x, out=lv)
        ops.nilif(x, out=lv, d_max=4.0, norm=4.0)
        ops.nilif(x, scale=sc, shift=sh, out=lv)
        ops.gemm_tc(a, packed, n=64, H=32, W=32, Cin=512, Cout=2048, scale=gsc, shift=gsh, pieces=3, want_spike=True)
        ops.linear_attn(q, k, v, n=64, Nq=1024, Nk=1024, heads=8, d=64, out_scale=64 ** -0.5 / 512)
        ops.peak_mma("i8", 1024)
else:
    cfg = getattr(s2f.configs, name)()
    seg = s2f.build_segmentor(cfg)
    seg.load_state_dict(synth.synthetic_checkpoint(name, cfg), strict=True)
    seg = seg.cuda()
    H, W = (512, 512) if name == "ade20k" else (1024, 2048)
    u8 = torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8).cuda()

    def run():
        with torch.no_grad():
            engine.segmentor_logits(seg, u8, labels=labels)

for _ in range(2):
    run()
torch.cuda.synchronize()
top = len(sys.argv) > 4 and sys.argv[4].startswith("top")
if top:
    # `top[K]`: only the first launch of each of the K most expensive (kernel class, layer shape) groups is profiled --
    # a `--set full` capture of all ~290 launches of a forward takes > 20 minutes under ncu, this one about one
    K = int(sys.argv[4][3:] or 16)
    prof = ops.Profiler()
    torch.cuda._sleep(int(3e8))
    ops.set_profiler(prof)
    run()
    ops.set_profiler(None)
    torch.cuda.synchronize()
    agg, first = {}, {}
    for i, (cls, e0, e1, _fl, _by, detail, _ex) in enumerate(prof.records):
        agg[(cls, detail)] = agg.get((cls, detail), 0.0) + e0.elapsed_time(e1)
        first.setdefault((cls, detail), i)
    chosen = sorted(agg, key=lambda k: -agg[k])[:K]
    want = {first[k]: k for k in chosen}
    print("profiled launches:", [(i, k, round(agg[k], 3)) for i, k in sorted(want.items())])
    state = {"i": -1, "on": False}

    def p0():
        state["i"] += 1
        if state["i"] in want:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            state["on"] = True
        return None

    def p1(e0, *a, **k):
        if state["on"]:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            state["on"] = False

    ops._p0, ops._p1 = p0, p1
    run()
    torch.cuda.synchronize()
else:
    torch.cuda.profiler.start()
    run()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("launches through the library so far:", ops.launch_count())
