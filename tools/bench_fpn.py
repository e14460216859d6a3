"""Micro-benchmark of the fused FPN merge launch (lateral 1x1 + BN + bilinear x2 of the coarser level + NI-LIF).
    python tools/bench_fpn.py [batch] [iters]        GPU box only"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spike2former_b200 import ops  # noqa: E402
from tools.bench_kernels import timed  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
IT = int(sys.argv[2]) if len(sys.argv) > 2 else 10
g = torch.Generator().manual_seed(0)
for (H, cin) in ((256, 32), (128, 64), (64, 128)):
    cout = 256
    a = torch.randint(0, 9, (B, H, H, cin), generator=g, dtype=torch.int8).cuda()
    w = torch.randn(cout, cin, generator=g) / cin ** 0.5
    sc, sh = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) + 1
    prev = (torch.randn(B, H // 2, H // 2, cout, generator=g) * 2).cuda()
    packed, rowscale = ops.pack_weights_i8(w, 1, cin, 3)
    kw = dict(n=B, H=H, W=H, Cin=cin, Cout=cout, scale=(sc * rowscale / 8).cuda(), shift=sh.cuda())
    pk = packed.cuda()
    t_plain = timed(lambda: ops.gemm_tc(a, pk, want_spike=True, **kw), IT)
    t_up = timed(lambda: ops.gemm_tc(a, pk, want_spike=True, up_prev=prev, **kw), IT)
    byt = a.numel() + prev.numel() * 4 + B * H * H * cout
    t_f16 = None
    if cin <= 64:
        w64 = torch.zeros(cout, 64)
        w64[:, :cin] = w
        pf, rs16 = ops.pack_pw_f16(w64)
        pf, sc16 = pf.cuda(), (sc.double() * rs16.double() / 8).float().cuda()
        shc = sh.cuda()
        t_f16 = timed(lambda: ops.fpn_merge_f16(a, pf, prev, n=B, H=H, W=H, Cin=cin, Cout=cout, scale=sc16, shift=shc), IT)
    print(f"{H}x{H} {cin}->{cout} B={B}: spike-only {t_plain * 1e6:8.1f} us | fused FPN merge {t_up * 1e6:8.1f} us "
          f"= {byt / t_up / 1e9:6.0f} GB/s algorithmic ({byt / 1e6:.0f} MB)"
          + (f" | fp16 kernel {t_f16 * 1e6:8.1f} us = {byt / t_f16 / 1e9:6.0f} GB/s" if t_f16 else ""))
