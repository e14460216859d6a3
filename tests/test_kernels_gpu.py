"""GPU parity of every C-ABI kernel against the oracle / a plain fp32 torch restatement, on seeded inputs."""
import pytest
import torch
import torch.nn.functional as F

from oracle import port
from spike2former_b200 import ops

pytestmark = pytest.mark.gpu


def gen(seed=0):
    return torch.Generator().manual_seed(seed)


def tie_vector():
    """Every rounding tie inside the clamp range plus the values around the clamps."""
    ties = [k + 0.5 for k in range(9)]
    eps = [0.5 - 2 ** -24, 0.5 + 2 ** -23, 7.5 - 2 ** -21, 7.5 + 2 ** -21]
    edge = [-0.0, 0.0, -1.0, 8.0, 8.25, 8.5, 9.0, 100.0, -100.0, 1e-30, 7.999999]
    return torch.tensor(ties + eps + edge, dtype=torch.float32)


# --------------------------------------------------------------------------------------------- NI-LIF
@pytest.mark.parametrize("shape", [(16,), (4, 64, 256), (3, 1000), (1, 7, 360), (2, 33, 33, 20), (0,)])
def test_nilif_bit_exact(shape):
    x = torch.rand(shape, generator=gen(1)) * 12 - 2                       # U(-2, 10): both clamps
    if x.numel() >= 32:
        tv = tie_vector()
        x.view(-1)[: tv.numel()] = tv
    want = torch.round(torch.clamp(0.0 + x, 0, 8)).to(torch.int8)          # Q_IFNode after reset
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    got, _, y = ops.nilif(x.cuda(), want_norm=True, C_=max(1, shape[-1]), ties=cnt.view(torch.int64))
    assert torch.equal(got.cpu(), want)
    assert torch.equal(y.cpu(), want.float() / 8)                          # the reference's "/ 8" output
    assert int(cnt.item()) == port.count_ties(x)


@pytest.mark.parametrize("d_max,norm", [(4.0, 4.0), (8.0, 8.0), (4.0, 8.0)])
def test_nilif_d4_variant_end_to_end(d_max, norm):
    """The D = 4 neuron of BASELINE.json config 2 (Quant4 / Multispike_norm: surrogate.py:541-557,
    mmseg/models/utils/Qtrick.py:4-38: round(clamp(x, 0, 4)) / 4) through the same kernel: levels, normalised
    output, tie count, membrane over T and the STE window [0, D]."""
    g = gen(41)
    x = torch.rand(64, 1024, 512 // 8, generator=g) * 7 - 1.5
    tv = torch.tensor([k + 0.5 for k in range(5)] + [4.0, 4.25, 4.5, 3.9999998, -0.0])
    x.view(-1)[: tv.numel()] = tv
    want = torch.round(torch.clamp(x, 0, d_max))
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    got, _, y = ops.nilif(x.cuda(), want_norm=True, d_max=d_max, norm=norm, ties=cnt)
    assert torch.equal(got.cpu(), want.to(torch.int8)) and torch.equal(y.cpu(), want / norm)
    assert int(cnt.item()) == port.count_ties(x, d_max) and int(got.max()) == int(min(d_max, 5))
    xs = (torch.rand(3, 4096, generator=g) - 0.3) * 5
    lv, v = port.nilif_reference(xs, d_max=d_max, T=3)
    got, vo, _ = ops.nilif(xs.cuda(), want_v_out=True, T=3, C_=64, d_max=d_max, norm=norm)
    assert torch.equal(got.cpu(), lv) and torch.equal(vo.cpu(), v)
    gy = torch.randn(4096, generator=g)
    gx = ops.nilif_bwd(xs[0].contiguous().cuda(), gy.cuda(), C_=64, d_max=d_max, norm=norm).cpu()
    assert torch.equal(gx, gy / norm * ((xs[0] >= 0) & (xs[0] <= d_max)).float())


def test_nilif_known_answers():
    """SURVEY.md section 0.4 [probed]: Q_IFNode on [0.5,1.5,2.5,3.5,7.5] gives levels [0,2,2,4,8]."""
    x = torch.tensor([0.5, 1.5, 2.5, 3.5, 7.5] + [0.0] * 11)
    got, _, _ = ops.nilif(x.cuda())
    assert got.cpu()[:5].tolist() == [0, 2, 2, 4, 8]


@pytest.mark.parametrize("C,N", [(64, 64 * 50), (360, 360 * 9), (20, 20 * 7)])
@pytest.mark.parametrize("d_max", [8.0, 4.0])
def test_nilif_affine_residual(C, N, d_max):
    g = gen(2)
    x = torch.rand(N, generator=g) * 12 - 2
    sc = torch.rand(C, generator=g) * 1.5 + 0.5
    sh = torch.rand(C, generator=g) * 2 - 1
    res = torch.randn(N, generator=g)
    u = (x.view(-1, C) * sc + sh).view(-1) + res
    want = torch.round(torch.clamp(u, 0, d_max)).to(torch.int8)
    got, _, _ = ops.nilif(x.cuda(), scale=sc.cuda(), shift=sh.cuda(), residual=res.cuda(), C_=C, d_max=d_max)
    assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("T,shape", [(4, (65, 15, 511)), (3, (8, 256)), (2, (5, 7))])
def test_nilif_multistep_membrane(T, shape):
    """Multi-step generalisation: membrane carried in registers over T, soft reset (neuron.py:133-153).
    Input pattern of neuron_kernel.py:1264-1310: (rand - 0.5) * 3, scaled up to reach several levels."""
    x = (torch.rand((T,) + shape, generator=gen(3)) - 0.5) * 6
    v0 = torch.rand(shape, generator=gen(4))
    want, v_want = port.nilif_reference(x, v0=v0, T=T)
    got, v, _ = ops.nilif(x.cuda(), v_in=v0.cuda(), want_v_out=True, T=T, C_=shape[-1])
    assert torch.equal(got.cpu(), want)
    assert torch.equal(v.cpu(), v_want)
    # stateful across calls == one longer call (MemoryModule semantics, base.py:25-52)
    g1, v1, _ = ops.nilif(x[:1].cuda(), v_in=v0.cuda(), want_v_out=True, T=1, C_=shape[-1])
    g2, v2, _ = ops.nilif(x[1:].cuda(), v_in=v1, want_v_out=True, T=T - 1, C_=shape[-1])
    assert torch.equal(torch.cat([g1, g2]).cpu(), want) and torch.equal(v2.cpu().view_as(v_want), v_want)


def test_nilif_positional_broadcast_and_transpose():
    g = gen(5)
    n, N, C = 3, 40, 32
    x = torch.randn(n, N, C, generator=g) * 3
    pos = torch.randn(N, C, generator=g)
    lvl = torch.randn(C, generator=g)
    want = torch.round(torch.clamp((x + lvl) + pos, 0, 8)).to(torch.int8)
    got, _, _ = ops.nilif(x.cuda(), scale=torch.ones(C).cuda(), shift=lvl.cuda(), residual=pos.cuda(),
                          residual_period=pos.numel())
    assert torch.equal(got.cpu(), want)
    # MSDA_FFN's reinterpreting reshape (mmcv_spike/transformer.py:777): [n,nq,C] read as [n,C,nq], stored transposed
    s = torch.round(torch.clamp(x, 0, 8)).to(torch.int8)
    want_t = s.reshape(n, C, N).permute(0, 2, 1).contiguous()
    got_t, _, _ = ops.nilif(x.cuda(), transpose=(N, C))
    assert torch.equal(got_t.cpu().view(n, N, C), want_t)


@pytest.mark.parametrize("shape,period", [((2, 16, 16, 64), 16 * 16 * 64), ((3, 100, 256), 100 * 256), ((1, 128, 128, 256), 128 * 128 * 256)])
def test_nilif_pair_equals_two_single_launches(shape, period):
    """One read of x, two neurons (with / without the positional residual): bit-equal to two s2f_nilif_fwd launches."""
    g = gen(31)
    x = (torch.rand(shape, generator=g) * 12 - 2).cuda()
    C = shape[-1]
    sc, sh = (torch.rand(C, generator=g) + 0.5).cuda(), torch.randn(C, generator=g).cuda()
    res = torch.randn(period, generator=g).cuda()
    a, b = ops.nilif_pair(x, sc, sh, res, residual_period=period)
    a1, _, _ = ops.nilif(x, scale=sc, shift=sh, residual=res, residual_period=period)
    b1, _, _ = ops.nilif(x, scale=sc, shift=sh)
    assert torch.equal(a, a1) and torch.equal(b, b1)


def test_nilif_backward_ste():
    """quant.backward (surrogate.py:531-538) through the reference's own autograd graph."""
    g = gen(6)
    x = (torch.rand(4096, generator=g) * 12 - 2).requires_grad_(True)
    x.data[:4] = torch.tensor([0.0, 8.0, -1e-6, 8.000001])
    gy = torch.randn(4096, generator=g)

    class Q(torch.autograd.Function):            # restatement of `quant`
        @staticmethod
        def forward(ctx, i):
            ctx.save_for_backward(i)
            return torch.round(torch.clamp(i, 0, 8))

        @staticmethod
        def backward(ctx, go):
            (i,) = ctx.saved_tensors
            gi = go.clone()
            gi[i < 0] = 0
            gi[i > 8] = 0
            return gi

    (Q.apply(x) / 8).backward(gy)
    got = ops.nilif_bwd(x.detach().cuda(), gy.cuda(), C_=1)
    assert torch.equal(got.cpu(), x.grad)


def test_q_ifnode_module_is_a_drop_in_with_surrogate_gradient():
    """spike2former_b200.Q_IFNode: forward == the reference neuron (values in {0, 1/8, .., 1}), membrane carried until
    reset() (MemoryModule semantics), backward == quant.backward / 8 (STE), all through the C ABI."""
    import spike2former_b200 as s2f

    g = gen(41)
    x = (torch.rand(4, 33, 64, generator=g) * 12 - 2)
    x.view(-1)[:9] = torch.tensor([0.5, 1.5, 2.5, 3.5, 4.5, 5.5, 6.5, 7.5, 8.5])
    node = s2f.Q_IFNode()
    xc = x.cuda().requires_grad_(True)
    y = node(xc)
    lv, v = port.nilif_reference(torch.stack([x, x]), T=2)               # oracle: two steps with the membrane carried
    assert torch.equal(y.detach().cpu(), lv[0].float() / 8)
    gy = torch.randn(x.shape, generator=g)
    y.backward(gy.cuda())
    want = gy / 8
    want[(x < 0) | (x > 8)] = 0
    assert torch.equal(xc.grad.cpu(), want)
    y2 = node(x.cuda())                                                  # second call: v from the first one
    assert torch.equal(y2.cpu(), lv[1].float() / 8)
    assert torch.equal(node.v.cpu(), v)
    node.reset()
    assert node.v == 0.0
    assert torch.equal(node(x.cuda()).cpu(), lv[0].float() / 8)
    with pytest.raises(RuntimeError):
        node(x)                                                          # CPU tensor: no CPU path


# --------------------------------------------------------------------------------------------- conv / linear
def _levels(shape, g, hi=9):
    return torch.randint(0, hi, shape, generator=g, dtype=torch.int8)


@pytest.mark.parametrize("cin,cout,k,stride,H,W", [(32, 64, 1, 1, 16, 16), (16, 48, 3, 1, 12, 20), (32, 64, 3, 2, 16, 16),
                                                   (360, 100, 1, 1, 8, 8), (3, 32, 7, 2, 32, 32), (20, 20, 1, 1, 5, 1)])
def test_conv_simt_vs_torch(cin, cout, k, stride, H, W):
    g = gen(7)
    n = 2
    real_input = cin == 3
    a = torch.randn(n, H, W, cin, generator=g) if real_input else _levels((n, H, W, cin), g)
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    sc = torch.rand(cout, generator=g) + 0.5
    sh = torch.randn(cout, generator=g)
    pad = (k - 1) // 2
    xin = (a.float() if real_input else a.float() / 8).permute(0, 3, 1, 2)
    ref = F.conv2d(xin.double(), w.double(), stride=stride, padding=pad) * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1)
    Ho, Wo = ref.shape[-2:]
    res = torch.randn(n, Ho, Wo, cout, generator=g)
    ref = ref.permute(0, 2, 3, 1) + res.double()
    w2d = ops.pad_rows4(w.permute(0, 2, 3, 1).reshape(cout, -1).contiguous().cuda())
    of, os_ = ops.conv_simt(a.cuda(), w2d, n=n, H=H, W=W, Cin=cin, Cout=cout, k=k, stride=stride, pad=pad,
                            scale=sc.cuda(), shift=sh.cuda(), residual=res.cuda(), want_f32=True, want_spike=True)
    err = (of.cpu().double() - ref).abs().max().item()
    assert err < 2e-5 * max(1.0, ref.abs().max().item()), err
    # spikes are the NI-LIF of the kernel's own fp32 output, bit for bit
    assert torch.equal(os_.cpu(), torch.round(torch.clamp(of.cpu(), 0, 8)).to(torch.int8))
    # transposed (channel-major) store
    oft, _ = ops.conv_simt(a.cuda(), w2d, n=n, H=H, W=W, Cin=cin, Cout=cout, k=k, stride=stride, pad=pad,
                           scale=sc.cuda(), shift=sh.cuda(), want_f32=True, transposed=True)
    of2, _ = ops.conv_simt(a.cuda(), w2d, n=n, H=H, W=W, Cin=cin, Cout=cout, k=k, stride=stride, pad=pad,
                           scale=sc.cuda(), shift=sh.cuda(), want_f32=True)
    assert torch.equal(oft.cpu().view(n, cout, Ho * Wo).permute(0, 2, 1), of2.cpu().view(n, Ho * Wo, cout))


@pytest.mark.parametrize("cin,cout,H,W", [(64, 32, 24, 20), (128, 64, 16, 16), (256, 128, 9, 7), (64, 16, 8, 8)])
def test_pointwise_fp32_3xtf32_vs_float64(cin, cout, H, W):
    """SepConv.pwconv2 path (sdtv2.py:176-178): fp32 activations x fp32 weights on the tensor cores with 3xTF32
    compensation must stay fp32-grade (the plain-TF32 error would be ~5e-4)."""
    g = gen(21)
    n = 3
    a = torch.randn(n, H, W, cin, generator=g) * 2
    w = torch.randn(cout, cin, generator=g) / cin ** 0.5
    sc, sh = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g)
    res = torch.randn(n, H, W, cout, generator=g)
    ref = (a.double().view(-1, cin) @ w.double().t()).view(n, H, W, cout) * sc.double() + sh.double() + res.double()
    of, os_ = ops.conv_simt(a.cuda(), ops.pad_rows4(w.cuda()), n=n, H=H, W=W, Cin=cin, Cout=cout, scale=sc.cuda(),
                            shift=sh.cuda(), residual=res.cuda(), want_f32=True, want_spike=True)
    err = (of.cpu().double() - ref).abs().max().item()
    assert err < 3e-6 * max(1.0, ref.abs().max().item()), err
    assert torch.equal(os_.cpu(), torch.round(torch.clamp(of.cpu(), 0, 8)).to(torch.int8))


@pytest.mark.parametrize("k", [3, 5, 7])
@pytest.mark.parametrize("spike_in", [True, False])
def test_dwconv_vs_torch(k, spike_in):
    g = gen(8)
    n, H, W, C = 2, 19, 23, 24
    a = _levels((n, H, W, C), g) if spike_in else torch.randn(n, H, W, C, generator=g)
    w = torch.randn(C, 1, k, k, generator=g) / k
    sc, sh = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    xin = (a.float() / 8 if spike_in else a).permute(0, 3, 1, 2)
    ref = F.conv2d(xin.double(), w.double(), padding=(k - 1) // 2, groups=C) * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1)
    w_tap = w.reshape(C, k * k).t().contiguous().cuda()
    of, os_ = ops.dwconv(a.cuda(), w_tap, n=n, H=H, W=W, C_=C, k=k, scale=sc.cuda(), shift=sh.cuda(), want_f32=True,
                         want_spike=True)
    assert (of.cpu().double() - ref.permute(0, 2, 3, 1)).abs().max().item() < 2e-5
    assert torch.equal(os_.cpu(), torch.round(torch.clamp(of.cpu(), 0, 8)).to(torch.int8))


@pytest.mark.parametrize("k,C,H,W", [(3, 256, 20, 37), (5, 512, 9, 16), (7, 64, 33, 40), (7, 24, 12, 12)])
def test_dwconv_spike_operands_bitwise_equal_to_fp32_operands(k, C, H, W):
    """The int8 path feeds the levels to the FFMAs as fp32 denormals against 2^100-scaled weights; every rounding must
    be the one the plain fp32 sum of w * (level / 8) makes, so both operand types give identical bits."""
    g = gen(21)
    n = 2
    a = _levels((n, H, W, C), g)
    w_tap = (torch.randn(k * k, C, generator=g) / k).cuda()
    sc, sh = (torch.rand(C, generator=g) + 0.5).cuda(), torch.randn(C, generator=g).cuda()
    f_i, s_i = ops.dwconv(a.cuda(), w_tap, n=n, H=H, W=W, C_=C, k=k, scale=sc, shift=sh, want_f32=True, want_spike=True)
    f_f, s_f = ops.dwconv((a.float() / 8).cuda(), w_tap, n=n, H=H, W=W, C_=C, k=k, scale=sc, shift=sh, want_f32=True,
                          want_spike=True)
    assert torch.equal(f_i, f_f)
    assert torch.equal(s_i, s_f)


@pytest.mark.parametrize("n,Nq,Nk,heads,d", [(2, 64, 64, 4, 16), (1, 20, 300, 4, 16), (2, 100, 1024, 8, 32), (1, 64, 64, 8, 45),
                                             (2, 1024, 1024, 8, 64), (1, 100, 4096, 8, 64),      # d = 64: BASELINE config 2
                                             (2, 1024, 1024, 8, 48),       # stage 4: heads straddle the 128-channel slabs (256-wide key box)
                                             (1, 100, 16384, 8, 32),       # batch 1: the keys are split over CTAs that add into kv
                                             (3, 100, 100, 8, 32), (2, 33, 777, 2, 8), (1, 50, 130, 5, 48)])   # ragged token tiles, narrow / odd head counts
def test_linear_attn_exact(n, Nq, Nk, heads, d):
    """(Q K^T) V == Q (K^T V) on integer levels; compared with the reference's op order in float64."""
    g = gen(9)
    C = heads * d
    q, k, v = _levels((n, Nq, C), g), _levels((n, Nk, C), g), _levels((n, Nk, C), g)
    scale = d ** -0.5 / 512
    hs = lambda t, N: t.double().view(n, N, heads, d).permute(0, 2, 1, 3)
    kv = hs(k, Nk).transpose(-2, -1) @ hs(v, Nk)
    ref = (hs(q, Nq) @ kv).permute(0, 2, 1, 3).reshape(n, Nq, C) * scale
    os_, of = ops.linear_attn(q.cuda(), k.cuda(), v.cuda(), n=n, Nq=Nq, Nk=Nk, heads=heads, d=d, out_scale=scale,
                              want_f32=True)
    assert (of.cpu().double() - ref).abs().max().item() <= 1e-6 * ref.abs().max().item()
    assert torch.equal(os_.cpu(), torch.round(torch.clamp(of.cpu(), 0, 8)).to(torch.int8))


def test_linear_attn_tensor_core_kv_equals_the_mma_sync_kernel(monkeypatch):
    """csrc/attn_tc.cu (K^T V on tcgen05, MN-major TMA operands) and the mma.sync kernel it replaces are both exact integer
    sums: their kv workspaces and outputs must be identical bit for bit, also for column slices of a fused q|k|v buffer."""
    import subprocess, sys, os
    g = gen(23)
    n, N, heads, d = 2, 640, 8, 48
    C = heads * d
    qkv = _levels((n, N, 3 * C), g).cuda()
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    os_, of = ops.linear_attn(q, k, v, n=n, Nq=N, Nk=N, heads=heads, d=d, out_scale=1e-4, q_ld=3 * C, kv_ld=3 * C, want_f32=True)
    hs = lambda t: t.double().cpu().reshape(n, N, heads, d).permute(0, 2, 1, 3)
    ref = (hs(q) @ (hs(k).transpose(-2, -1) @ hs(v))).permute(0, 2, 1, 3).reshape(n, N, C) * 1e-4
    assert (of.cpu().double() - ref).abs().max().item() <= 1e-6 * ref.abs().max().item()
    # spikes only: the per-image spike GEMM then takes its spike-only epilogue instead of the staged fp32 one
    os2, _ = ops.linear_attn(q, k, v, n=n, Nq=N, Nk=N, heads=heads, d=d, out_scale=1e-4, q_ld=3 * C, kv_ld=3 * C)
    assert torch.equal(os2, os_)
    # the legacy kernel in a fresh process (the switch is read once per process)
    code = ("import torch, sys; sys.path.insert(0, %r); from spike2former_b200 import ops; qkv = torch.load(sys.argv[1]).cuda();"
            "C = %d; q, k, v = qkv[..., :C], qkv[..., C:2*C], qkv[..., 2*C:];"
            "s, f = ops.linear_attn(q, k, v, n=%d, Nq=%d, Nk=%d, heads=%d, d=%d, out_scale=1e-4, q_ld=3*C, kv_ld=3*C, want_f32=True);"
            "torch.save((s.cpu(), f.cpu()), sys.argv[2])") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), C, n, N, N, heads, d)
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        torch.save(qkv.cpu(), os.path.join(td, "in.pt"))
        env = dict(os.environ, S2F_ATTN_TC="0")
        subprocess.check_call([sys.executable, "-c", code, os.path.join(td, "in.pt"), os.path.join(td, "out.pt")], env=env)
        s_old, f_old = torch.load(os.path.join(td, "out.pt"))
    assert torch.equal(of.cpu(), f_old) and torch.equal(os_.cpu(), s_old)


@pytest.mark.parametrize("n,Nq,Nk,heads,d,masked", [(2, 100, 1024, 8, 32, True), (1, 37, 300, 4, 45, True), (2, 20, 64, 2, 64, True),
                                                    (1, 100, 4096, 8, 32, False)])
def test_dec_attn_masked_matches_reference_order(n, Nq, Nk, heads, d, masked):
    """s2f_dec_attn: scores = Q K^T / sqrt(dim); masked_fill(mask, 0); out = scores V; NI-LIF
    (mmcv_spike/transformer.py:262-270, 345-353) -- against the same expression in float64, and against the linear
    kernel when there is no mask."""
    g = gen(17)
    C = heads * d
    q, k, v = _levels((n, Nq, C), g), _levels((n, Nk, C), g), _levels((n, Nk, C), g)
    mask = (torch.rand(n * heads, Nq, Nk, generator=g) < 0.3) if masked else None
    scale = 1.0 / (C ** 0.5) / 512
    hs = lambda t, N: t.double().view(n, N, heads, d).permute(0, 2, 1, 3)
    scores = hs(q, Nq) @ hs(k, Nk).transpose(-2, -1)
    if masked:
        scores = scores.masked_fill(mask.view(n, heads, Nq, Nk), 0)
    ref = (scores @ hs(v, Nk)).permute(0, 2, 1, 3).reshape(n, Nq, C) * scale
    os_, of = ops.dec_attn(q.cuda(), k.cuda(), v.cuda(), n=n, Nq=Nq, Nk=Nk, heads=heads, d=d, out_scale=scale,
                           mask=mask.cuda() if masked else None, want_f32=True)
    assert (of.cpu().double() - ref).abs().max().item() <= 1e-6 * ref.abs().max().item()
    assert torch.equal(os_.cpu(), torch.round(torch.clamp(of.cpu(), 0, 8)).to(torch.int8))
    if not masked:
        ls, lf = ops.linear_attn(q.cuda(), k.cuda(), v.cuda(), n=n, Nq=Nq, Nk=Nk, heads=heads, d=d, out_scale=scale, want_f32=True)
        assert torch.equal(ls, os_) and torch.equal(lf, of)


def test_decoder_layer_with_cross_attention_mask_vs_oracle():
    """engine.head_forward(cross_attn_masks=...) against the oracle's decoder with the same masks
    (detr_layers.py:491-559 `cross_attn_mask`): query states after every layer."""
    import spike2former_b200 as s2f
    from oracle import port, weights
    from spike2former_b200 import engine

    cfg = s2f.configs.tiny()
    P = weights.calibrated_state(cfg, 64, 64)
    img = weights.test_image(cfg, 64, 64, batch=2)
    heads = cfg["decode_head"]["transformer_decoder"]["layer_cfg"]["self_attn_cfg"]["num_heads"]
    nq = cfg["decode_head"]["num_queries"]
    g = gen(23)
    masks = [torch.rand(2 * heads, nq, (64 // 16) ** 2 * 4 ** (i % 3), generator=g) < 0.4 for i in range(6)]
    # oracle: the port's decoder_layer takes cross_mask
    cx = port.Ctx(P)
    cx.marks = {}
    with torch.no_grad():
        feats = port.backbone_forward(cx, cfg["backbone"], img)
        orig = port.decoder_layer
        calls = []

        def with_mask(cx_, key, query, kv, qpos, kpos, heads_, cross_mask=None):
            i = int(key.rsplit(".", 1)[1])
            calls.append(i)
            return orig(cx_, key, query, kv, qpos, kpos, heads_, masks[i])

        port.decoder_layer = with_mask
        try:
            port.head_forward(cx, cfg["decode_head"], feats, 1)
        finally:
            port.decoder_layer = orig
    assert calls == list(range(6))
    seg = s2f.build_segmentor(cfg)
    seg.load_state_dict(P, strict=True)
    seg = seg.cuda()

    class Rec(engine.NullProbe):
        observe = True

        def __init__(self, prefix="", store=None):
            self.prefix, self.store = prefix, ({} if store is None else store)

        def scoped(self, p):
            return Rec(self.prefix + p, self.store)

        def real(self, name, t, layout="cm"):
            self.store[self.prefix + name] = t.detach().clone()
            return t

    rec = Rec("decode_head.")
    with torch.no_grad():
        f2 = engine.backbone_forward(seg.backbone, img.cuda())
        engine.head_forward(seg.decode_head, f2, rec, cross_attn_masks=[m.cuda() for m in masks])
    for i in range(6):
        key = f"decode_head.transformer_decoder.layers.{i}.out"
        want, got = cx.marks[key], rec.store[key].cpu()
        assert (got - want).abs().max() <= 1e-3 * want.abs().max(), (i, float((got - want).abs().max()))
    # and the masks matter
    rec2 = Rec("decode_head.")
    with torch.no_grad():
        engine.head_forward(seg.decode_head, f2, rec2)
    assert not torch.equal(rec2.store["decode_head.transformer_decoder.layers.5.out"], rec.store["decode_head.transformer_decoder.layers.5.out"])


def test_linear_attn_strided_operands_and_padded_output():
    """q|k|v as column slices of one fused [n, N, 3C] projection (engine._ms_block) and a 16-byte padded output row."""
    g = gen(12)
    n, N, heads, d = 2, 200, 8, 32
    C = heads * d
    qkv = _levels((n, N, 3 * C), g).cuda()
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    scale = d ** -0.5 / 512
    hs = lambda t: t.cpu().double().view(n, N, heads, d).permute(0, 2, 1, 3)
    ref = (hs(q) @ (hs(k).transpose(-2, -1) @ hs(v))).permute(0, 2, 1, 3).reshape(n, N, C) * scale
    os_, of = ops.linear_attn(q, k, v, n=n, Nq=N, Nk=N, heads=heads, d=d, q_ld=3 * C, kv_ld=3 * C, out_ld=C + 16,
                              out_scale=scale, want_f32=True)
    assert (of.cpu()[..., :C].double() - ref).abs().max().item() <= 1e-6 * ref.abs().max().item()
    assert torch.equal(os_.cpu()[..., :C], torch.round(torch.clamp(of.cpu()[..., :C], 0, 8)).to(torch.int8))
    assert int(os_.cpu()[..., C:].abs().sum()) == 0 and float(of.cpu()[..., C:].abs().sum()) == 0.0


def test_linear_attn_long_keys_need_wide_accumulation():
    """Nk large enough that 8 * 64 * Nk * d exceeds int32: the 64-bit accumulation path (Cityscapes level 2)."""
    g = gen(13)
    n, Nq, Nk, heads, d = 1, 40, 140000, 2, 32
    C = heads * d
    q, k, v = _levels((n, Nq, C), g), _levels((n, Nk, C), g), _levels((n, Nk, C), g)
    hs = lambda t, N: t.double().view(n, N, heads, d).permute(0, 2, 1, 3)
    ref = (hs(q, Nq) @ (hs(k, Nk).transpose(-2, -1) @ hs(v, Nk))).permute(0, 2, 1, 3).reshape(n, Nq, C) * 1e-9
    _, of = ops.linear_attn(q.cuda(), k.cuda(), v.cuda(), n=n, Nq=Nq, Nk=Nk, heads=heads, d=d, out_scale=1e-9, want_f32=True)
    assert (of.cpu().double() - ref).abs().max().item() <= 1e-6 * ref.abs().max().item()


@pytest.mark.parametrize("n,H,W,G,Cg", [(2, 8, 8, 4, 16), (1, 32, 32, 32, 8), (2, 5, 9, 8, 8)])
def test_dcnv3_gather_vs_reference_core(n, H, W, G, Cg):
    """Pattern of ops_dcnv3/test.py:33-60 (seed 3, inputs*0.01, offsets*10, K=3, pad 1), fp32 tolerance of that
    file (rtol 1e-2 / atol 1e-3) tightened to 1e-5; the oracle is the port of dcnv3_core_pytorch."""
    g = gen(3)
    K = 3
    x = torch.rand(n, H, W, G * Cg, generator=g) * 0.01
    off = torch.rand(n, H, W, G * K * K * 2, generator=g) * 10 - 5
    mask = _levels((n, H, W, G * K * K), g)
    for os_ in (1.0, 2.0):
        ref = port.dcnv3_core(x, off, mask.float() / 8, K, 1, 1, 1, G, Cg, os_)
        got = ops.dcnv3_gather(x.cuda(), off.cuda(), mask.cuda(), n=n, H=H, W=W, G=G, Cg=Cg, K=K, offset_scale=os_)
        assert (got.cpu() - ref).abs().max().item() < 1e-5 * max(ref.abs().max().item(), 1e-3)


def test_upsample_add_lif_vs_torch():
    g = gen(10)
    n, C, Hp, Wp = 2, 16, 6, 10
    prev = torch.randn(n, Hp, Wp, C, generator=g) * 3
    cur = torch.randn(n, 2 * Hp, 2 * Wp, C, generator=g) * 3
    up = F.interpolate(prev.permute(0, 3, 1, 2), size=(2 * Hp, 2 * Wp), mode="bilinear", align_corners=False)
    ref = cur + up.permute(0, 2, 3, 1)
    sp, f = ops.upsample_add_lif(cur.cuda(), prev.cuda(), n=n, H=2 * Hp, W=2 * Wp, Hp=Hp, Wp=Wp, C_=C, want_f32=True)
    assert (f.cpu() - ref).abs().max().item() < 1e-5
    assert torch.equal(sp.cpu(), torch.round(torch.clamp(f.cpu(), 0, 8)).to(torch.int8))


def test_semantic_tail_vs_torch():
    g = gen(11)
    n, Q, K, h, w = 2, 20, 11, 12, 16
    mp = torch.randn(n, Q, h, w, generator=g) * 3
    cls = torch.randn(n, Q, K + 1, generator=g) * 4
    up = F.interpolate(mp, size=(2 * h, 2 * w), mode="bilinear", align_corners=False)
    ref = torch.einsum("bqc,bqhw->bchw", F.softmax(cls, -1)[..., :-1], up.sigmoid())
    mpp = mp.permute(0, 2, 3, 1).contiguous().cuda()
    got = ops.semantic_tail_simt(mpp, cls.cuda(), n=n, Q=Q, K=K, h=h, w=w, H=2 * h, W=2 * w)
    assert (got.cpu() - ref).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())
    # tensor-core tail: bf16 hi+lo split operands keep ~16 mantissa bits
    got_tc, lab = ops.semantic_tail(mpp, cls.cuda(), n=n, Q=Q, K=K, h=h, w=w, H=2 * h, W=2 * w, want_labels=True)
    assert (got_tc.cpu() - ref).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item())
    assert torch.equal(lab.cpu().long(), got_tc.cpu().argmax(1))
    assert (lab.cpu().long() == ref.argmax(1)).float().mean().item() > 0.999


@pytest.mark.parametrize("n,Q,K,h,w", [(2, 100, 150, 64, 64), (1, 100, 19, 20, 24), (3, 37, 150, 9, 50), (1, 128, 256, 16, 16)])
def test_semantic_tail_pipelined_x2(n, Q, K, h, w):
    """The warp-specialised x2 tail (tail_x2_kernel): ADE20K / Cityscapes class counts, ragged tiles, odd Q."""
    g = gen(14)
    mp = torch.randn(n, Q, h, w, generator=g) * 3
    cls = torch.randn(n, Q, K + 1, generator=g) * 4
    up = F.interpolate(mp, size=(2 * h, 2 * w), mode="bilinear", align_corners=False)
    ref = torch.einsum("bqc,bqhw->bchw", F.softmax(cls, -1)[..., :-1], up.sigmoid())
    mpp = mp.permute(0, 2, 3, 1).contiguous().cuda()
    got, lab = ops.semantic_tail(mpp, cls.cuda(), n=n, Q=Q, K=K, h=h, w=w, H=2 * h, W=2 * w, want_labels=True)
    assert (got.cpu() - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    assert torch.equal(lab.cpu().long(), got.cpu().argmax(1))
    assert (lab.cpu().long() == ref.argmax(1)).float().mean().item() > 0.999
    _, lab2 = ops.semantic_tail(mpp, cls.cuda(), n=n, Q=Q, K=K, h=h, w=w, H=2 * h, W=2 * w, want_logits=False, want_labels=True)
    assert torch.equal(lab2.cpu(), lab.cpu())


def test_semantic_tail_general_scale_falls_back_to_serial_kernel():
    g = gen(15)
    n, Q, K, h, w, H, W = 1, 20, 11, 12, 16, 30, 50
    mp = torch.randn(n, Q, h, w, generator=g) * 3
    cls = torch.randn(n, Q, K + 1, generator=g) * 4
    up = F.interpolate(mp, size=(H, W), mode="bilinear", align_corners=False)
    ref = torch.einsum("bqc,bqhw->bchw", F.softmax(cls, -1)[..., :-1], up.sigmoid())
    got, _ = ops.semantic_tail(mp.permute(0, 2, 3, 1).contiguous().cuda(), cls.cuda(), n=n, Q=Q, K=K, h=h, w=w, H=H, W=W)
    assert (got.cpu() - ref).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item())


def test_cpu_tensors_are_rejected():
    with pytest.raises(RuntimeError):
        ops.nilif(torch.zeros(16))


@pytest.mark.parametrize("k,Cm,Cout,H,W,n", [(7, 64, 32, 40, 48, 2), (7, 128, 64, 19, 23, 2), (7, 256, 128, 16, 32, 3),
                                             (3, 64, 16, 9, 17, 1)])
def test_sepconv_dwpw_fused_vs_float64_and_two_kernel_path(k, Cm, Cout, H, W, n):
    """SepConv tail (sdtv2.py:176-179) in one launch: depthwise stencil (reference tap order) + pwconv2 on
    tcgen05 kind::f16 with fp16 hi/lo operands.  Against float64 and against s2f_dwconv + the fp32 1x1 kernel."""
    g = gen(31)
    a = _levels((n, H, W, Cm), g)
    w_dw = torch.randn(Cm, 1, k, k, generator=g) / k
    w_pw = torch.randn(Cout, Cm, generator=g) / Cm ** 0.5
    w_pw[0, :] *= 1e-3                                            # a tiny row and a heavy-tailed row
    w_pw[1, 3] = 20.0
    sc, sh = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g)
    res = torch.randn(n, H, W, Cout, generator=g)
    x2 = F.conv2d((a.double() / 8).permute(0, 3, 1, 2), w_dw.double(), padding=(k - 1) // 2, groups=Cm)
    ref = torch.einsum("nchw,oc->nhwo", x2, w_pw.double()) * sc.double() + sh.double() + res.double()
    w_tap = w_dw.reshape(Cm, k * k).t().contiguous().cuda()
    packed, rowscale = ops.pack_pw_f16(w_pw)
    a_pre = 16.0
    scale = (sc.double() * rowscale.double() / a_pre).float().cuda()
    of, os_ = ops.sepconv_dwpw(a.cuda(), w_tap, packed.cuda(), n=n, H=H, W=W, Cm=Cm, Cout=Cout, k=k, scale=scale,
                               shift=sh.cuda(), a_pre=a_pre, residual=res.cuda(), want_f32=True, want_spike=True)
    tol = 2e-5 * max(1.0, ref.abs().max().item())
    err = (of.cpu().double() - ref).abs().max().item()
    assert err < tol, (err, tol)
    # levels: equal to the oracle's except where its pre-activation is within the tolerance of a rounding boundary
    lv_ref = torch.round(torch.clamp(ref, 0, 8))
    bad = os_.cpu().double() != lv_ref
    frac = (torch.clamp(ref, 0, 8) - torch.floor(torch.clamp(ref, 0, 8)) - 0.5).abs()
    assert bool((frac[bad] < 1e-4 * ref.abs().clamp_min(1.0)[bad]).all()), int(bad.sum())
    assert torch.equal(os_.cpu(), torch.round(torch.clamp(of.cpu(), 0, 8)).to(torch.int8))
    # the two-kernel path it replaces
    d, _ = ops.dwconv(a.cuda(), w_tap, n=n, H=H, W=W, C_=Cm, k=k, want_f32=True)
    o2, _ = ops.conv_simt(d, ops.pad_rows4(w_pw.cuda()), n=n, H=H, W=W, Cin=Cm, Cout=Cout, scale=sc.cuda(), shift=sh.cuda(),
                          residual=res.cuda(), want_f32=True, want_spike=True)
    assert (of - o2).abs().max().item() < tol


@pytest.mark.parametrize("H,W,cin,n", [(32, 48, 32, 2), (24, 32, 64, 2), (18, 22, 32, 1), (256, 256, 32, 1)])
def test_fpn_merge_f16_vs_float64_and_digit_plane_kernel(H, W, cin, n):
    """Finest FPN merges (pixel_decoder.py:451-462) on csrc/fpn_tc.cu: lateral 1x1 on tcgen05 kind::f16 (fp16 hi / lo
    weights, levels exact in fp16) + bilinear x2 + NI-LIF, against float64 and the int8 digit-plane kernel's fused merge."""
    g = gen(41)
    cout = 256
    a = _levels((n, H, W, cin), g)
    w = torch.randn(cout, cin, generator=g) / cin ** 0.5
    w[3] *= 1e-3
    w[5, 1] = 9.0
    sc, sh = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) + 1
    prev = torch.randn(n, H // 2, W // 2, cout, generator=g) * 2
    lat = torch.einsum("nhwc,oc->nhwo", a.double() / 8, w.double()) * sc.double() + sh.double()
    up = F.interpolate(prev.double().permute(0, 3, 1, 2), size=(H, W), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    ref = lat + up
    w64 = torch.zeros(cout, 64)
    w64[:, :cin] = w
    packed, rowscale = ops.pack_pw_f16(w64)
    scale = (sc.double() * rowscale.double() / 8).float().cuda()
    _, lv = ops.fpn_merge_f16(a.cuda(), packed.cuda(), prev.cuda(), n=n, H=H, W=W, Cin=cin, Cout=cout, scale=scale, shift=sh.cuda())
    lv_ref = torch.round(torch.clamp(ref, 0, 8))
    bad = lv.cpu().double() != lv_ref
    frac = (torch.clamp(ref, 0, 8) - torch.floor(torch.clamp(ref, 0, 8)) - 0.5).abs()
    assert bool((frac[bad] < 1e-4 * ref.abs().clamp_min(1.0)[bad]).all()), int(bad.sum())
    assert int(bad.sum()) <= 2e-4 * bad.numel()
    # the digit-plane kernel's fused merge (exact integer lateral conv): same levels up to the same near-ties
    pk, rs = ops.pack_weights_i8(w, 1, cin, 3)
    _, lv2 = ops.gemm_tc(a.cuda(), pk.cuda(), n=n, H=H, W=W, Cin=cin, Cout=cout, scale=(sc * rs / 8).cuda(), shift=sh.cuda(),
                         want_spike=True, up_prev=prev.cuda())
    diff = lv != lv2
    assert bool((frac[diff.cpu()] < 1e-4 * ref.abs().clamp_min(1.0)[diff.cpu()]).all()), int(diff.sum())
