"""Micro-benchmarks of the individual kernels (GPU box only): device time by CUDA events, L2 flushed between launches.

    python tools/bench_kernels.py gemm            # spike GEMM shapes of the model + BASELINE config 2
    python tools/bench_kernels.py nilif           # fused NI-LIF at BASELINE config 2 size (HBM GB/s)
    python tools/bench_kernels.py all [--iters N]
Prints one line per case and a JSON summary (gpurun_out/kernels.json when that directory exists).
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spike2former_b200 import ops  # noqa: E402

ITERS = 10
for i, a in enumerate(sys.argv):
    if a == "--iters":
        ITERS = int(sys.argv[i + 1])
FLUSH = None


def flush_l2():
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")
    FLUSH.view(torch.int64).sum()        # read 256 MB: evicts the working set and leaves only clean lines in L2


def timed(fn, iters=ITERS, flush=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(400000)        # ~0.2 ms spin: the launch below is already queued when the GPU reaches e0
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3          # median seconds


GEMM_CASES = [
    # name, n, H, W, Cin, Cout, k, stride, f32, spike, residual
    ("cfg2 qkv-proj 1x1 C=512 B=64", 64, 32, 32, 512, 512, 1, 1, False, True, False),
    ("cfg2 mlp fc1 512->2048 B=64", 64, 32, 32, 512, 2048, 1, 1, False, True, False),
    ("cfg2 mlp fc2 2048->512 B=64", 64, 32, 32, 2048, 512, 1, 1, True, True, True),
    ("fc1 256->1024 @32^2", 16, 32, 32, 256, 1024, 1, 1, False, True, False),
    ("fc2 1024->256 @32^2", 16, 32, 32, 1024, 256, 1, 1, True, True, True),
    ("pw1 256->512 @32^2", 16, 32, 32, 256, 512, 1, 1, False, True, False),
    ("qkv 3x3 256->768 @32^2", 16, 32, 32, 256, 768, 3, 1, False, True, False),
    ("conv1 3x3 32->128 @256^2", 16, 256, 256, 32, 128, 3, 1, False, True, False),
    ("conv2 3x3 128->32 @256^2", 16, 256, 256, 128, 32, 3, 1, True, True, True),
    ("conv1 3x3 128->512 @64^2", 16, 64, 64, 128, 512, 3, 1, False, True, False),
    ("conv2 3x3 512->128 @64^2", 16, 64, 64, 512, 128, 3, 1, True, True, True),
    ("dec k/v 256->256 Nk=16384", 16, 16384, 1, 256, 256, 1, 1, False, True, False),
    ("lateral 32->256 @256^2 f32", 16, 256, 256, 32, 256, 1, 1, True, False, False),
    ("down 3x3s2 64->128 @128^2", 16, 128, 128, 64, 128, 3, 2, True, True, False),
]


def bench_gemm(out):
    g = torch.Generator().manual_seed(0)
    for name, n, H, W, cin, cout, k, stride, f32, spike, res in GEMM_CASES:
        a = torch.randint(0, 9, (n, H, W, cin), generator=g, dtype=torch.int8).cuda()
        w = torch.randn(cout, k * k * cin, generator=g) / (cin * k * k) ** 0.5
        PIECES = int(os.environ.get("PIECES", "3"))          # experiments: fewer digit planes
        packed, rowscale = ops.pack_weights_i8(w, k * k, cin, PIECES)
        packed = packed.cuda()
        sc, sh = (rowscale / 8).cuda(), torch.zeros(cout).cuda()
        pad = (k - 1) // 2
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        r = torch.randn(n, Ho, Wo, cout, device="cuda") if res else None

        def run():
            ops.gemm_tc(a, packed, n=n, H=H, W=W, Cin=cin, Cout=cout, scale=sc, shift=sh, k=k, stride=stride, pad=pad,
                        pieces=PIECES, residual=r, want_f32=f32, want_spike=spike)

        t = timed(run)
        flops = 2.0 * n * Ho * Wo * cout * k * k * cin
        nbytes = a.numel() + packed.numel() + n * Ho * Wo * cout * ((4 if f32 else 0) + (1 if spike else 0) + (4 if res else 0))
        line = dict(kernel="gemm_i8_tc", case=name, us=t * 1e6, tflops=flops / t / 1e12, gbs=nbytes / t / 1e9)
        out.append(line)
        print(f"{name:34s} {t * 1e6:9.1f} us  {flops / t / 1e12:8.1f} TFLOP/s (algorithmic)  {nbytes / t / 1e9:7.0f} GB/s")


def bench_nilif(out):
    g = torch.Generator().manual_seed(0)
    for name, shape, T in [("cfg2 NI-LIF B=64 C=512 N=1024 T=1", (64, 1024, 512), 1),
                           ("cfg2 NI-LIF T=4 (membrane in registers)", (4, 16, 1024, 512), 4),
                           ("NI-LIF 16x256x256x256", (16, 256, 256, 256), 1)]:
        x = (torch.rand(shape, generator=g) * 12 - 2).cuda()
        C = shape[-1]
        sc, sh = torch.ones(C).cuda(), torch.zeros(C).cuda()
        lv = torch.empty(shape, dtype=torch.int8, device="cuda")
        for aff in (False, True):
            def run():
                ops.nilif(x, scale=sc if aff else None, shift=sh if aff else None, T=T, C_=C, out=lv)

            t = timed(run)
            nbytes = x.numel() * 5
            line = dict(kernel="nilif", case=f"{name} affine={int(aff)}", us=t * 1e6, gbs=nbytes / t / 1e9)
            out.append(line)
            print(f"{line['case']:50s} {t * 1e6:9.1f} us  {nbytes / t / 1e9:7.0f} GB/s (algorithmic 5 B/neuron-step)")


DW_CASES = [
    # name, n, H, W, C, k, f32 out, spike out
    ("SepConv dw7 @256^2 x64", 32, 256, 256, 64, 7, True, False),
    ("SepConv dw7 @128^2 x128", 32, 128, 128, 128, 7, True, False),
    ("SepConv dw7 @64^2 x256", 32, 64, 64, 256, 7, True, False),
    ("FPN dw3 @256^2 x256", 32, 256, 256, 256, 3, False, True),
    ("FPN dw3 @128^2 x256 f32", 32, 128, 128, 256, 3, True, False),
    ("DCN dw5 @32^2 x512", 32, 32, 32, 512, 5, False, True),
    ("DCN dw5 @32^2 x256", 32, 32, 32, 256, 5, False, True),
    ("Conv dw3 @32^2 x512", 32, 32, 32, 512, 3, False, True),
]


def bench_dwconv(out):
    """Depthwise stencil: the two-row register-tiled kernel against the one-row kernel it replaced (S2F_DW_LEGACY=1)."""
    g = torch.Generator().manual_seed(0)
    for name, n, H, W, C, k, f32, spike in DW_CASES:
        a = torch.randint(0, 9, (n, H, W, C), generator=g, dtype=torch.int8).cuda()
        w = (torch.randn(k * k, C, generator=g) / k).cuda()
        sc, sh = (torch.rand(C, generator=g) + 0.5).cuda(), torch.randn(C, generator=g).cuda()

        def run():
            return ops.dwconv(a, w, n=n, H=H, W=W, C_=C, k=k, scale=sc, shift=sh, want_f32=f32, want_spike=spike)

        res = {}
        for mode in ("0", "1"):
            os.environ["S2F_DW_LEGACY"] = mode
            res[mode] = (timed(run), run())
        os.environ["S2F_DW_LEGACY"] = "0"
        same = all(torch.equal(x, y) for x, y in zip(res["0"][1], res["1"][1]) if x is not None)
        t = res["0"][0]
        nbytes = a.numel() * (1 + 4 * f32 + spike)
        fma = 2.0 * a.numel() * k * k
        line = dict(kernel="dwconv", case=name, us=t * 1e6, legacy_us=res["1"][0] * 1e6, gbs=nbytes / t / 1e9,
                    tflops=fma / t / 1e12, bitwise_equal_to_legacy=same)
        out.append(line)
        print(f"{name:34s} {t * 1e6:8.1f} us (legacy {res['1'][0] * 1e6:8.1f})  {nbytes / t / 1e9:6.0f} GB/s  "
              f"{fma / t / 1e12:5.1f} TFLOP/s fp32  bitwise==legacy {same}")


def bench_sepconv(out):
    """Fused SepConv tail (dw7x7 + pwconv2 + BN + residual + LIF) against the two launches it replaces."""
    g = torch.Generator().manual_seed(0)
    for name, n, H, W, Cm, Cout in [("SepConv @256^2 64->32", 32, 256, 256, 64, 32), ("SepConv @128^2 128->64", 32, 128, 128, 128, 64),
                                    ("SepConv @64^2 256->128", 32, 64, 64, 256, 128)]:
        a = torch.randint(0, 9, (n, H, W, Cm), generator=g, dtype=torch.int8).cuda()
        wd = (torch.randn(49, Cm, generator=g) / 7).cuda()
        wp = torch.randn(Cout, Cm, generator=g) / Cm ** 0.5
        packed, rowscale = ops.pack_pw_f16(wp)
        packed = packed.cuda()
        sc, sh = (torch.rand(Cout, generator=g) + 0.5), torch.randn(Cout, generator=g).cuda()
        scale = (sc.double() * rowscale.double() / 16.0).float().cuda()
        res = torch.randn(n, H, W, Cout, generator=g).cuda()

        def run():
            return ops.sepconv_dwpw(a, wd, packed, n=n, H=H, W=W, Cm=Cm, Cout=Cout, k=7, scale=scale, shift=sh, a_pre=16.0,
                                    residual=res, want_f32=True, want_spike=True)

        t = timed(run)
        nbytes = a.numel() + res.numel() * 9
        fl = 2.0 * n * H * W * (Cm * 49 + Cm * Cout)
        line = dict(kernel="sepconv_dwpw", case=name, us=t * 1e6, gbs=nbytes / t / 1e9, tflops=fl / t / 1e12)
        out.append(line)
        print(f"{name:34s} {t * 1e6:8.1f} us  {nbytes / t / 1e9:6.0f} GB/s  {fl / t / 1e12:5.1f} TFLOP/s")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    out = []
    if what in ("gemm", "all"):
        bench_gemm(out)
    if what in ("nilif", "all"):
        bench_nilif(out)
    if what in ("dwconv", "all"):
        bench_dwconv(out)
    if what in ("sepconv", "all"):
        bench_sepconv(out)
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        json.dump(out, open(os.path.join(d, f"kernels_{what}.json"), "w"), indent=1)
