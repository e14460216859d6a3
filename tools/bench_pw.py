"""Micro-benchmark of the fp32 pointwise (pwconv2) and stem kernels (GPU box only)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spike2former_b200 import ops
g = torch.Generator().manual_seed(0)
for n, H, W, cin, cout in [(32, 256, 256, 64, 32), (32, 128, 128, 128, 64), (32, 64, 64, 256, 128)]:
    a = torch.randn(n, H, W, cin, generator=g).cuda()
    w = ops.pad_rows4((torch.randn(cout, cin, generator=g) / cin ** 0.5).cuda())
    sc, sh = torch.ones(cout).cuda(), torch.zeros(cout).cuda()
    res = torch.randn(n, H, W, cout, device="cuda")
    def run():
        ops.conv_simt(a, w, n=n, H=H, W=W, Cin=cin, Cout=cout, scale=sc, shift=sh, residual=res, want_f32=True, want_spike=True)
    for _ in range(3): run()
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(400000); e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[2] * 1e-3
    nbytes = a.numel() * 4 + res.numel() * 9
    print(f"pw {cin}->{cout} @{H}x{W} x{n}: {t*1e6:.1f} us  {nbytes/t/1e9:.0f} GB/s  {2*n*H*W*cin*cout/t/1e12:.1f} TFLOP/s")
