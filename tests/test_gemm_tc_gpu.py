"""GPU parity of the tcgen05 int8 spike GEMM against a float64 restatement (and the CUDA-core kernel)."""
import pytest
import torch
import torch.nn.functional as F

from spike2former_b200 import ops

pytestmark = pytest.mark.gpu


def _case(cin, cout, k, stride, H, W, n=2, pieces=3, seed=0, transposed=False, residual=True):
    g = torch.Generator().manual_seed(seed)
    a = torch.randint(0, 9, (n, H, W, cin), generator=g, dtype=torch.int8)
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    w[0, 0, 0, 0] = 0.75          # exact power-of-two-ish row maxima exercise the packer's exponent choice
    sc = torch.rand(cout, generator=g) + 0.5
    sh = torch.randn(cout, generator=g)
    pad = (k - 1) // 2
    ref = F.conv2d((a.double() / 8).permute(0, 3, 1, 2), w.double(), stride=stride, padding=pad)
    ref = ref * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1)
    Ho, Wo = ref.shape[-2:]
    ref = ref.permute(0, 2, 3, 1)
    res = torch.randn(n, Ho, Wo, cout, generator=g) if residual else None
    if residual:
        ref = ref + res.double()
    w2d = w.permute(0, 2, 3, 1).reshape(cout, -1).contiguous()
    packed, rowscale = ops.pack_weights_i8(w2d, k * k, cin, pieces)
    scale_tc = (sc.double() * rowscale.double() / 8).float().cuda()
    of, os_ = ops.gemm_tc(a.cuda(), packed.cuda(), n=n, H=H, W=W, Cin=cin, Cout=cout, scale=scale_tc, shift=sh.cuda(),
                          k=k, stride=stride, pad=pad, pieces=pieces, residual=res.cuda() if residual else None,
                          want_f32=True, want_spike=True, transposed=transposed)
    torch.cuda.synchronize()
    of, os_ = of.cpu(), os_.cpu()
    if transposed:
        of = of.view(n, cout, Ho * Wo).permute(0, 2, 1).reshape(n, Ho, Wo, cout)
        os_ = os_.view(n, cout, Ho * Wo).permute(0, 2, 1).reshape(n, Ho, Wo, cout)
    return of, os_, ref


SHAPES = [
    # cin, cout, k, stride, H, W
    (32, 64, 1, 1, 16, 16),       # SWIZZLE_32B, exact tiles
    (64, 128, 1, 1, 16, 24),      # SWIZZLE_64B
    (128, 64, 1, 1, 8, 16),       # SWIZZLE_128B, one chunk
    (256, 256, 1, 1, 32, 32),     # stage-3 1x1
    (368, 100, 1, 1, 10, 10),     # ragged K chunk (368 = 2*128 + 112), ragged Cout, M tail (200 rows)
    (1024, 256, 1, 1, 32, 32),    # MLP fc2: 8 chunks
    (32, 128, 3, 1, 32, 32),      # ConvBlock1_1.conv1 shape class
    (128, 32, 3, 1, 16, 16),      # conv2: Cout 32 (half-empty N tile)
    (64, 64, 3, 1, 12, 20),       # spatial sizes that are not tile multiples
    (32, 64, 3, 2, 32, 32),       # strided downsample (TMA element strides)
    (128, 256, 3, 2, 16, 16),
    (256, 768, 3, 1, 32, 32),     # merged q|k|v RepConv
    (48, 80, 3, 1, 9, 7),         # odd everything
]


@pytest.mark.parametrize("cin,cout,k,stride,H,W", SHAPES)
def test_gemm_tc_matches_float64(cin, cout, k, stride, H, W):
    of, os_, ref = _case(cin, cout, k, stride, H, W)
    scale = max(1.0, ref.abs().max().item())
    err = (of.double() - ref).abs().max().item()
    assert err < 5e-6 * scale, (err, scale)
    assert torch.equal(os_, torch.round(torch.clamp(of, 0, 8)).to(torch.int8))


@pytest.mark.parametrize("pieces,tol", [(1, 3e-2), (2, 3e-4), (3, 5e-6)])
def test_gemm_tc_digit_planes(pieces, tol):
    of, _, ref = _case(256, 128, 1, 1, 16, 16, pieces=pieces)
    assert (of.double() - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item())


def test_gemm_tc_transposed_and_no_residual():
    of, os_, ref = _case(256, 288, 1, 1, 32, 32, transposed=True, residual=False)
    assert (of.double() - ref).abs().max().item() < 5e-6 * max(1.0, ref.abs().max().item())
    assert torch.equal(os_, torch.round(torch.clamp(of, 0, 8)).to(torch.int8))


@pytest.mark.parametrize("cin,cout,k,H,W", [(256, 256, 1, 32, 32), (368, 100, 1, 10, 10), (64, 64, 3, 12, 20), (32, 256, 1, 64, 64)])
def test_gemm_tc_epilogue_paths_agree(cin, cout, k, H, W):
    """Spike-only launches keep the row-per-lane epilogue, fp32 / residual launches go through the shared-memory
    transposed (coalesced) epilogue: same accumulators, same affine, so the levels must be identical."""
    g = torch.Generator().manual_seed(11)
    n = 3
    a = torch.randint(0, 9, (n, H, W, cin), generator=g, dtype=torch.int8).cuda()
    w = torch.randn(cout, k * k * cin, generator=g) / (cin * k * k) ** 0.5
    sc, sh = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) + 1
    packed, rowscale = ops.pack_weights_i8(w, k * k, cin, 3)
    kw = dict(n=n, H=H, W=W, Cin=cin, Cout=cout, scale=(sc * rowscale / 8).cuda(), shift=sh.cuda(), k=k, pad=(k - 1) // 2)
    _, s_only = ops.gemm_tc(a, packed.cuda(), want_spike=True, **kw)
    f_both, s_both = ops.gemm_tc(a, packed.cuda(), want_f32=True, want_spike=True, **kw)
    f_only, _ = ops.gemm_tc(a, packed.cuda(), want_f32=True, **kw)
    f_tr, s_tr = ops.gemm_tc(a, packed.cuda(), want_f32=True, want_spike=True, transposed=True, **kw)
    assert torch.equal(s_only, s_both)
    assert torch.equal(f_only, f_both)
    assert torch.equal(f_tr.view(n, cout, H * W).permute(0, 2, 1).reshape(n, H, W, cout), f_both)
    assert torch.equal(s_tr.view(n, cout, H * W).permute(0, 2, 1).reshape(n, H, W, cout), s_both)


def test_gemm_tc_matches_cuda_core_kernel_bitwise_spikes():
    """Same inputs through both kernels: fp32 outputs agree to fp32 rounding, spikes differ only at near-ties."""
    g = torch.Generator().manual_seed(5)
    n, H, W, cin, cout = 4, 32, 32, 256, 256
    a = torch.randint(0, 9, (n, H, W, cin), generator=g, dtype=torch.int8).cuda()
    w = (torch.randn(cout, cin, generator=g) / 16)
    sc, sh = (torch.rand(cout, generator=g) + 0.5), torch.randn(cout, generator=g) + 1
    packed, rowscale = ops.pack_weights_i8(w, 1, cin, 3)
    f_tc, s_tc = ops.gemm_tc(a, packed.cuda(), n=n, H=H, W=W, Cin=cin, Cout=cout, scale=(sc * rowscale / 8).cuda(),
                             shift=sh.cuda(), want_f32=True, want_spike=True)
    f_cc, s_cc = ops.conv_simt(a, ops.pad_rows4(w.cuda()), n=n, H=H, W=W, Cin=cin, Cout=cout, scale=sc.cuda(),
                               shift=sh.cuda(), want_f32=True, want_spike=True)
    assert (f_tc - f_cc).abs().max().item() < 1e-5
    diff = (s_tc != s_cc)
    frac = (f_cc - f_cc.floor() - 0.5).abs()
    assert int((diff & (frac > 1e-4)).sum()) == 0
    assert int(diff.sum()) <= 1e-4 * diff.numel()


@pytest.mark.parametrize("cin,H,W", [(32, 32, 32), (64, 16, 24), (128, 8, 8)])
def test_gemm_tc_fused_fpn_merge_matches_separate_kernels(cin, H, W):
    """lateral 1x1 + BN + bilinear_up(prev) + NI-LIF in one launch == gemm_tc -> s2f_upsample_add_lif up to the
    compiler's FMA contraction of the interpolation (pixel_decoder.py:451-462)."""
    g = torch.Generator().manual_seed(7)
    n, cout = 2, 256
    a = torch.randint(0, 9, (n, H, W, cin), generator=g, dtype=torch.int8).cuda()
    w = torch.randn(cout, cin, generator=g) / cin ** 0.5
    sc, sh = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) + 1
    prev = (torch.randn(n, H // 2, W // 2, cout, generator=g) * 2).cuda()
    packed, rowscale = ops.pack_weights_i8(w, 1, cin, 3)
    kw = dict(n=n, H=H, W=W, Cin=cin, Cout=cout, scale=(sc * rowscale / 8).cuda(), shift=sh.cuda())
    cur, _ = ops.gemm_tc(a, packed.cuda(), want_f32=True, **kw)
    want_s, want_f = ops.upsample_add_lif(cur, prev, n=n, H=H, W=W, Hp=H // 2, Wp=W // 2, C_=cout, want_f32=True)
    got_f, got_s = ops.gemm_tc(a, packed.cuda(), want_f32=True, want_spike=True, up_prev=prev, **kw)
    assert (got_f - want_f).abs().max().item() < 2e-6 * max(1.0, want_f.abs().max().item())
    assert torch.equal(got_s, torch.round(torch.clamp(got_f, 0, 8)).to(torch.int8))      # spikes of the kernel's own fp32
    flips = got_s != want_s
    assert int(flips.sum()) <= 1e-4 * flips.numel()
    assert int((flips & ((want_f - want_f.floor() - 0.5).abs() > 1e-4)).sum()) == 0      # only at rounding ties
    up = F.interpolate(prev.permute(0, 3, 1, 2), size=(H, W), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    assert (got_f - (cur + up)).abs().max().item() < 1e-5


@pytest.mark.parametrize("cin,cout,k", [(256, 256, 1), (128, 512, 3), (1024, 256, 1)])
def test_gemm_tc_heavy_tailed_rows(cin, cout, k):
    """Rows with a few large weights over a background ~1000x smaller (BN-folded checkpoints, and the stable synthetic
    init): the 21-bit fixed point is relative to the row maximum, so small weights keep only ~11 bits each -- the
    OUTPUT error must still be bounded by 2^-21 x row maximum per term, i.e. as tight as for a flat row."""
    g = torch.Generator().manual_seed(7)
    n, H, W = 2, 16, 16
    K = cin * k * k
    a = torch.randint(0, 9, (n, H, W, cin), generator=g, dtype=torch.int8)
    w2d = torch.randn(cout, K, generator=g) * (1e-3 / K ** 0.5)
    idx = torch.stack([torch.randperm(K, generator=g)[:3] for _ in range(cout)])
    w2d.scatter_add_(1, idx, torch.randn(cout, 3, generator=g))
    w2d[0] = torch.randn(K, generator=g).abs() * 1e-6                  # a row of tiny weights only
    w2d[1, 5] = 1000.0                                                 # one huge outlier
    sc, sh = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g)
    w4 = w2d.view(cout, k, k, cin).permute(0, 3, 1, 2)
    ref = F.conv2d((a.double() / 8).permute(0, 3, 1, 2), w4.double(), padding=(k - 1) // 2).permute(0, 2, 3, 1)
    ref = ref * sc.double() + sh.double()
    packed, rowscale = ops.pack_weights_i8(w2d, k * k, cin, 3)
    of, os_ = ops.gemm_tc(a.cuda(), packed.cuda(), n=n, H=H, W=W, Cin=cin, Cout=cout,
                          scale=(sc.double() * rowscale.double() / 8).float().cuda(), shift=sh.cuda(), k=k, pad=(k - 1) // 2,
                          want_f32=True, want_spike=True)
    rowmax = w2d.abs().amax(1).double()
    # per output: K terms x level <= 1 x half a quantisation step (< 2^-20 rowmax), plus fp32 rounding of the result
    bound = (K * 2.0 ** -20 * rowmax * sc.double()).view(1, 1, 1, -1) + 3e-7 * ref.abs().clamp(min=1.0)
    err = (of.cpu().double() - ref).abs()
    assert bool((err <= bound).all()), float((err / bound).max())
    assert torch.equal(os_.cpu(), torch.round(torch.clamp(of.cpu(), 0, 8)).to(torch.int8))
