"""Operator-level Python interface over the C ABI (include/s2f.h).

Each function takes torch CUDA tensors, checks dtype / device / contiguity, allocates the outputs
with torch (device memory + stream plumbing only) and launches the kernels on the current stream.
CPU tensors are rejected: there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ConvArgs, GemmTcArgs, S2FError, check

D_MAX = 8.0     # Quant() clamp max: surrogate.py:526
NORM = 8.0      # hard-coded "/ 8": neuron.py:197


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, dtype=None, name="tensor"):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise S2FError(f"{name}: CUDA tensor required (no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise S2FError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise S2FError(f"{name}: must be contiguous")
    return C.c_void_p(t.data_ptr())


def launch_count() -> int:
    return int(_lib.lib().s2f_launch_count())


class Profiler:
    """Per-launch CUDA-event timing on the launch stream, grouped by kernel class (bench.py's roofline)."""

    def __init__(self):
        self.records = []          # (class, ev0, ev1, alg_flops, alg_bytes)

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for cls, e0, e1, fl, by, _, ex in self.records:
            a = agg.setdefault(cls, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0, exec_flops=0.0))
            a["ms"] += e0.elapsed_time(e1); a["flops"] += fl; a["bytes"] += by; a["launches"] += 1; a["exec_flops"] += ex
        return agg


_PROF = None


def set_profiler(p):
    global _PROF
    _PROF = p


def _p0():
    if _PROF is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _p1(e0, cls, flops=0.0, nbytes=0.0, detail="", exec_flops=None):
    """flops: ALGORITHMIC work of the reference layer(s) this launch stands for; exec_flops: what the kernel executes."""
    if e0 is None:
        return
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    _PROF.records.append((cls, e0, e1, float(flops), float(nbytes), detail, float(flops if exec_flops is None else exec_flops)))


def _nb(*ts):
    return sum(t.numel() * t.element_size() for t in ts if t is not None)


# ------------------------------------------------------------------------------------------ NI-LIF
def nilif(x, scale=None, shift=None, residual=None, residual_period=0, v_in=None, want_v_out=False, want_norm=False,
          T=1, C_=None, d_max=D_MAX, norm=NORM, transpose=None, ties=None, out=None):
    """Fused NI-LIF.  x: fp32 [T, ...] (T leading when T > 1).  Returns (levels int8 same shape, v_out|None, y|None).

    The reference call it replaces: `Q_IFNode(surrogate_function=Quant())(x)` -- neuron.py:462-550."""
    _ptr(x, torch.float32, "x")
    total = x.numel()
    if total % T != 0:
        raise S2FError("nilif: numel not divisible by T")
    N = total // T
    C_ = int(C_ if C_ is not None else x.shape[-1])
    levels = out if out is not None else torch.empty(x.shape, dtype=torch.int8, device=x.device)
    v_out = torch.empty(x.shape[1:] if T > 1 else x.shape, dtype=torch.float32, device=x.device) if want_v_out else None
    y = torch.empty_like(x) if want_norm else None
    tr, tc = (transpose if transpose else (0, 0))
    e0 = _p0()
    check(_lib.lib().s2f_nilif_fwd(_ptr(x), _ptr(scale, torch.float32, "scale"), _ptr(shift, torch.float32, "shift"),
                                   _ptr(residual, torch.float32, "residual"), int(residual_period),
                                   _ptr(v_in, torch.float32, "v_in"), _ptr(v_out), _ptr(levels, torch.int8, "levels"),
                                   _ptr(y), int(T), int(N), C_, float(d_max), float(norm), int(tr), int(tc),
                                   _ptr(ties), _stream()), "s2f_nilif_fwd")
    _p1(e0, "nilif", 0, _nb(x, levels, v_in, v_out, y) + (x.numel() * 4 if residual is not None else 0),
        f"{tuple(x.shape)} aff={int(scale is not None)} res={int(residual is not None)} tr={int(bool(transpose))}")
    return levels, v_out, y


def nilif_pair(x, scale, shift, residual, residual_period=0, C_=None, d_max=D_MAX):
    """(NI-LIF(x*scale + shift + residual), NI-LIF(x*scale + shift)) from one read of x -- the decoder's key / value
    inputs of a pyramid level.  Returns two int8 level tensors shaped like x."""
    _ptr(x, torch.float32, "x")
    C_ = int(C_ if C_ is not None else x.shape[-1])
    a = torch.empty(x.shape, dtype=torch.int8, device=x.device)
    b = torch.empty(x.shape, dtype=torch.int8, device=x.device)
    e0 = _p0()
    check(_lib.lib().s2f_nilif_pair(_ptr(x), _ptr(scale, torch.float32, "scale"), _ptr(shift, torch.float32, "shift"),
                                    _ptr(residual, torch.float32, "residual"), int(residual_period), _ptr(a), _ptr(b),
                                    int(x.numel()), C_, float(d_max), _stream()), "s2f_nilif_pair")
    _p1(e0, "nilif", 0, _nb(x, a, b), f"{tuple(x.shape)} pair")
    return a, b


def nilif_bwd(x, gy, scale=None, shift=None, residual=None, C_=None, d_max=D_MAX, norm=NORM):
    """Surrogate gradient of the neuron (T=1): quant.backward, surrogate.py:531-538."""
    gx = torch.empty_like(x)
    check(_lib.lib().s2f_nilif_bwd(_ptr(x, torch.float32, "x"), _ptr(scale), _ptr(shift), _ptr(residual),
                                   _ptr(gy, torch.float32, "gy"), _ptr(gx), x.numel(),
                                   int(C_ if C_ is not None else x.shape[-1]), float(d_max), float(norm), _stream()),
          "s2f_nilif_bwd")
    return gx


def nilif_train_fwd(x, d_max=D_MAX, norm=NORM):
    """Training-mode neuron forward: (y = level / norm fp32, tag uint8 = level | 0x80 where x is outside [0, d_max])."""
    _ptr(x, torch.float32, "x")
    y = torch.empty_like(x)
    tag = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    check(_lib.lib().s2f_nilif_train_fwd(_ptr(x), _ptr(y), _ptr(tag), x.numel(), float(d_max), float(norm), _stream()),
          "s2f_nilif_train_fwd")
    return y, tag


def nilif_train_bwd(tag, gy, norm=NORM):
    """Surrogate gradient from the saved tag byte: gx = gy / norm where the neuron's input was inside [0, d_max]."""
    _ptr(tag, torch.uint8, "tag")
    _ptr(gy, torch.float32, "gy")
    gx = torch.empty_like(gy)
    check(_lib.lib().s2f_nilif_train_bwd(_ptr(tag), _ptr(gy), _ptr(gx), gy.numel(), float(norm), _stream()), "s2f_nilif_train_bwd")
    return gx


def affine_add_lif(x, scale=None, residual=None, want_f32=True, want_spike=True, d_max=D_MAX):
    out_f = torch.empty_like(x) if want_f32 else None
    out_s = torch.empty(x.shape, dtype=torch.int8, device=x.device) if want_spike else None
    e0 = _p0()
    check(_lib.lib().s2f_affine_add_lif(_ptr(x, torch.float32, "x"), _ptr(scale, torch.float32, "scale"),
                                        _ptr(residual, torch.float32, "residual"), _ptr(out_f), _ptr(out_s),
                                        x.numel(), int(x.shape[-1]), float(d_max), _stream()), "s2f_affine_add_lif")
    _p1(e0, "elementwise", 0, _nb(x, residual, out_f, out_s))
    return out_f, out_s


# ------------------------------------------------------------------------------------------ conv / linear
def pad_rows4(w2d: torch.Tensor) -> torch.Tensor:
    """[Cout, K] -> [Cout, ceil4(K)] zero padded (row layout s2f_conv_simt expects)."""
    k = w2d.shape[1]
    kp = (k + 3) // 4 * 4
    if kp == k:
        return w2d.contiguous()
    out = torch.zeros(w2d.shape[0], kp, dtype=w2d.dtype, device=w2d.device)
    out[:, :k] = w2d
    return out


def conv_simt(a, w, *, n, H, W, Cin, Cout, k=1, stride=1, pad=0, scale=None, shift=None, residual=None,
              a_scale=1.0 / NORM, want_f32=False, want_spike=False, transposed=False, a_img_stride=0, a_stride_m=0,
              a_stride_k=0, w_img_stride=0, d_max=D_MAX, out_f32=None, out_spike=None):
    """General implicit-GEMM convolution (CUDA cores).  a: int8 levels or fp32, channels-last."""
    is_spike = a.dtype == torch.int8
    if not is_spike and a.dtype != torch.float32:
        raise S2FError("conv_simt: a must be int8 levels or fp32")
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    shape = (n, Cout, Ho * Wo) if transposed else (n, Ho, Wo, Cout)
    if want_f32 and out_f32 is None:
        out_f32 = torch.empty(shape, dtype=torch.float32, device=a.device)
    if want_spike and out_spike is None:
        out_spike = torch.empty(shape, dtype=torch.int8, device=a.device)
    args = ConvArgs()
    args.a, args.a_is_spike, args.a_scale = _ptr(a, name="a"), int(is_spike), float(a_scale)
    args.w, args.scale, args.shift = _ptr(w, torch.float32, "w"), _ptr(scale, torch.float32, "scale"), _ptr(shift, torch.float32, "shift")
    args.residual = _ptr(residual, torch.float32, "residual")
    args.out_f32, args.out_spike, args.out_transposed = _ptr(out_f32), _ptr(out_spike), int(transposed)
    args.n, args.H, args.W, args.Cin, args.Cout = n, H, W, Cin, Cout
    args.KH = args.KW = k
    args.stride, args.pad = stride, pad
    args.a_img_stride, args.a_stride_m, args.a_stride_k = int(a_img_stride), int(a_stride_m), int(a_stride_k)
    args.w_img_stride, args.d_max = int(w_img_stride), float(d_max)
    e0 = _p0()
    check(_lib.lib().s2f_conv_simt(C.byref(args), _stream()), "s2f_conv_simt")
    _p1(e0, "gemm_simt", 2.0 * n * Ho * Wo * Cout * k * k * Cin, _nb(a, residual, out_f32, out_spike) + Cout * k * k * Cin * 4,
        f"{n}x{H}x{W} {Cin}->{Cout} k{k}s{stride} {a.dtype}")
    return out_f32, out_spike


def pack_weights_i8(w2d: torch.Tensor, taps: int, cin: int, pieces: int = 3):
    """Host-side: fp32 [Cout, taps*cin] (cin fastest) -> (int8 digit planes in the kernel's tile layout, rowscale[Cout])."""
    w = w2d.detach().to("cpu", torch.float32).contiguous()
    cout = w.shape[0]
    if w.shape[1] != taps * cin:
        raise S2FError("pack_weights_i8: weight row length != taps * cin")
    lib = _lib.lib()
    nbytes = lib.s2f_pack_weights_i8(None, cout, taps, cin, pieces, None, None)
    if nbytes <= 0:
        raise S2FError("pack_weights_i8: bad arguments")
    packed = torch.empty(nbytes, dtype=torch.int8)
    rowscale = torch.empty(cout, dtype=torch.float32)
    got = lib.s2f_pack_weights_i8(C.c_void_p(w.data_ptr()), cout, taps, cin, pieces, C.c_void_p(packed.data_ptr()),
                                  C.c_void_p(rowscale.data_ptr()))
    if got != nbytes:
        raise S2FError("pack_weights_i8 failed")
    return packed, rowscale


def tc_eligible(cin: int, k: int, stride: int) -> bool:
    return cin >= 32 and cin % 16 == 0 and k in (1, 3) and stride in (1, 2)


def pack_rows_i8_device(w, *, n_img, rows_per_img, K, pieces=3, post_scale=1.0, bias_col=None):
    """Device-side packer for per-image weight matrices: w fp32 [n_img*rows_per_img, ld] on the GPU ->
    (digit planes in the per-image tile layout, scale [rows], shift [rows])."""
    _ptr(w, torch.float32, "w")
    ld = w.shape[1]
    tiles_n = (rows_per_img + 63) // 64
    packed = torch.zeros(n_img * tiles_n * pieces * 64 * K, dtype=torch.int8, device=w.device)
    sc = torch.empty(n_img * rows_per_img, dtype=torch.float32, device=w.device)
    sh = torch.empty_like(sc)
    bias_ptr = C.c_void_p(w.data_ptr() + 4 * bias_col) if bias_col is not None else None
    e0 = _p0()
    check(_lib.lib().s2f_pack_rows_i8_device(_ptr(w), ld, n_img, rows_per_img, K, pieces, _ptr(packed), _ptr(sc), _ptr(sh),
                                             bias_ptr, float(post_scale), _stream()), "s2f_pack_rows_i8_device")
    _p1(e0, "elementwise", 0, _nb(w, packed))
    return packed, sc, sh


def gemm_tc(a, w_packed, *, n, H, W, Cin, Cout, scale, shift, k=1, stride=1, pad=0, pieces=3, residual=None,
            want_f32=False, want_spike=False, transposed=False, d_max=D_MAX, per_image=False, up_prev=None, alg_macs=None):
    """tcgen05 spike GEMM.  a: int8 levels channels-last; `scale` already contains rowscale * 1/8."""
    if a.dtype != torch.int8:
        raise S2FError("gemm_tc: a must be int8 levels")
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    shape = (n, Cout, Ho * Wo) if transposed else (n, Ho, Wo, Cout)
    out_f32 = torch.empty(shape, dtype=torch.float32, device=a.device) if want_f32 else None
    out_spike = torch.empty(shape, dtype=torch.int8, device=a.device) if want_spike else None
    args = GemmTcArgs()
    args.a, args.w_packed = _ptr(a, torch.int8, "a"), _ptr(w_packed, torch.int8, "w_packed")
    args.scale, args.shift = _ptr(scale, torch.float32, "scale"), _ptr(shift, torch.float32, "shift")
    args.residual = _ptr(residual, torch.float32, "residual")
    args.out_f32, args.out_spike, args.out_transposed = _ptr(out_f32), _ptr(out_spike), int(transposed)
    args.n, args.H, args.W, args.Cin, args.Cout = n, H, W, Cin, Cout
    args.KH = args.KW = k
    args.stride, args.pad, args.pieces, args.d_max = stride, pad, pieces, float(d_max)
    args.per_image_weights = int(per_image)
    if up_prev is not None:                      # fused top-down FPN merge: y += bilinear_up(up_prev)
        if up_prev.dim() != 4 or up_prev.shape[0] != n or up_prev.shape[3] != Cout:
            raise S2FError("gemm_tc: up_prev must be fp32 [n, Hp, Wp, Cout]")
        args.up_prev, args.up_H, args.up_W = _ptr(up_prev, torch.float32, "up_prev"), int(up_prev.shape[1]), int(up_prev.shape[2])
    e0 = _p0()
    check(_lib.lib().s2f_gemm_i8_tc(C.byref(args), _stream()), "s2f_gemm_i8_tc")
    executed = 2.0 * n * Ho * Wo * Cout * k * k * Cin
    _p1(e0, "gemm_tc", executed if alg_macs is None else 2.0 * n * Ho * Wo * alg_macs,
        _nb(a, residual, out_f32, out_spike) + Cout * k * k * Cin * pieces,
        f"{n}x{H}x{W} {Cin}->{Cout} k{k}s{stride} f32={int(want_f32)} sp={int(want_spike)} res={int(residual is not None)} tr={int(transposed)}",
        exec_flops=executed)
    return out_f32, out_spike


def dwconv(a, w_tap, *, n, H, W, C_, k, scale=None, shift=None, a_scale=1.0 / NORM, want_f32=False, want_spike=False,
           no_pad=False, d_max=D_MAX):
    """Depthwise k x k (stride 1).  w_tap: fp32 [k*k, C]."""
    is_spike = a.dtype == torch.int8
    pad = 0 if no_pad else (k - 1) // 2
    Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    out_f = torch.empty((n, Ho, Wo, C_), dtype=torch.float32, device=a.device) if want_f32 else None
    out_s = torch.empty((n, Ho, Wo, C_), dtype=torch.int8, device=a.device) if want_spike else None
    e0 = _p0()
    check(_lib.lib().s2f_dwconv(_ptr(a, name="a"), int(is_spike), float(a_scale), _ptr(w_tap, torch.float32, "w_tap"),
                                _ptr(scale, torch.float32, "scale"), _ptr(shift, torch.float32, "shift"), None,
                                _ptr(out_f), _ptr(out_s), n, H, W, C_, k, int(no_pad), float(d_max), _stream()),
          "s2f_dwconv")
    _p1(e0, "dwconv", 2.0 * n * Ho * Wo * C_ * k * k, _nb(a, out_f, out_s),
        f"{n}x{H}x{W}x{C_} k{k} f32={int(want_f32)} sp={int(want_spike)}")
    return out_f, out_s


def pack_pw_f16(w2d: torch.Tensor):
    """fp32 [Cout, Cm] 1x1 weights -> (uint8 image, rowscale) for s2f_sepconv_dwpw: every row is scaled by the power of
    two that brings its largest magnitude into [1024, 2048) and split into fp16 hi + lo (22 significant bits of the row
    maximum, like the 21-bit digit planes of the spike GEMM); image = [hi | lo][Cm / 64][Np rows][64 k] in the K-major
    SWIZZLE_128B order of tl_off (csrc/tail_tc.cu).  True weight = (hi + lo) * rowscale."""
    w = w2d.detach().double().cpu()
    cout, cm = w.shape
    if cm % 64:
        raise S2FError("pack_pw_f16: Cm must be a multiple of 64")
    npad = (cout + 15) // 16 * 16
    amax = w.abs().amax(dim=1).clamp_min(1e-30)
    e = torch.floor(torch.log2(amax))                       # amax in [2^e, 2^(e+1))
    rowscale = torch.pow(torch.tensor(2.0, dtype=torch.float64), e - 10)
    ws = w / rowscale[:, None]
    hi = ws.to(torch.float16)
    lo = (ws - hi.double()).to(torch.float16)
    img = torch.zeros(2, cm // 64, npad, 64, dtype=torch.float16)
    r = torch.arange(cout)
    for plane, src in enumerate((hi, lo)):
        blocks = src.reshape(cout, cm // 64, 8, 8)           # [row, 128-byte atom, 16-byte chunk, element]
        for chunk in range(8):
            idx = ((chunk ^ (r & 7))[:, None] * 8 + torch.arange(8)[None, :])   # swizzle: chunk ^ (row % 8)
            for a in range(cm // 64):
                img[plane, a].index_put_((r[:, None].expand(-1, 8), idx), blocks[:, a, chunk, :])
    return img.view(torch.uint8).reshape(-1).contiguous(), rowscale.to(torch.float32)


def sepconv_dwpw(a, w_dw_tap, w_pw_packed, *, n, H, W, Cm, Cout, k, scale, shift, a_pre=16.0, residual=None,
                 a_scale=1.0 / NORM, want_f32=True, want_spike=True, d_max=D_MAX, alg_macs=None):
    """SepConv tail (sdtv2.py:176-179) in one launch: depthwise k x k over int8 levels, then pwconv2 + BN (+ residual).
    `scale` must already hold rowscale / a_pre (see pack_pw_f16)."""
    if a.dtype != torch.int8:
        raise S2FError("sepconv_dwpw: a must be int8 levels")
    out_f = torch.empty((n, H, W, Cout), dtype=torch.float32, device=a.device) if want_f32 else None
    out_s = torch.empty((n, H, W, Cout), dtype=torch.int8, device=a.device) if want_spike else None
    need = int(_lib.lib().s2f_sepconv_bpack_bytes(Cm, Cout))
    if w_pw_packed.numel() * w_pw_packed.element_size() != need:
        raise S2FError(f"sepconv_dwpw: packed weights hold {w_pw_packed.numel()} bytes, expected {need}")
    e0 = _p0()
    check(_lib.lib().s2f_sepconv_dwpw(_ptr(a, torch.int8, "a"), float(a_scale), _ptr(w_dw_tap, torch.float32, "w_dw"),
                                      _ptr(w_pw_packed, torch.uint8, "w_pw_packed"), float(a_pre),
                                      _ptr(scale, torch.float32, "scale"), _ptr(shift, torch.float32, "shift"),
                                      _ptr(residual, torch.float32, "residual"), _ptr(out_f), _ptr(out_s),
                                      n, H, W, Cm, Cout, k, float(d_max), _stream()), "s2f_sepconv_dwpw")
    macs = Cm * k * k + (Cm * Cout if alg_macs is None else alg_macs)
    _p1(e0, "sepconv_dwpw", 2.0 * n * H * W * macs, _nb(a, residual, out_f, out_s), f"{n}x{H}x{W} dw{k} {Cm} -> pw {Cout}")
    return out_f, out_s


def fpn_merge_f16(a, w_packed, prev, *, n, H, W, Cin, Cout, scale, shift, d_max=D_MAX, alg_macs=None):
    """Lateral 1x1 + BN + bilinear x2 of the coarser level + NI-LIF in one launch (pixel_decoder.py:451-462) for
    Cin <= 64, Cout = 256: w_packed = pack_pw_f16(W zero-padded to 64 input channels), scale = rowscale * a_scale * BN."""
    if a.dtype != torch.int8:
        raise S2FError("fpn_merge_f16: a must be int8 levels")
    Hp, Wp = int(prev.shape[1]), int(prev.shape[2])
    out_s = torch.empty((n, H, W, Cout), dtype=torch.int8, device=a.device)
    e0 = _p0()
    check(_lib.lib().s2f_fpn_merge_f16(_ptr(a, torch.int8, "a"), _ptr(w_packed, torch.uint8, "w_packed"),
                                       _ptr(scale, torch.float32, "scale"), _ptr(shift, torch.float32, "shift"),
                                       _ptr(prev, torch.float32, "prev"), _ptr(out_s), n, H, W, Cin, Cout, Hp, Wp,
                                       float(d_max), _stream()), "s2f_fpn_merge_f16")
    _p1(e0, "fpn_merge", 2.0 * n * H * W * (Cin * Cout if alg_macs is None else alg_macs), _nb(a, prev, out_s),
        f"{n}x{H}x{W} {Cin}->{Cout} + up x2")
    return None, out_s


# ------------------------------------------------------------------------------------------ attention / DCN / tail
def linear_attn(q, k, v, *, n, Nq, Nk, heads, d, out_scale, q_ld=None, kv_ld=None, out_ld=None, want_f32=False,
                d_max=D_MAX):
    """Spike-driven attention without softmax: NI-LIF((Q (K^T V)) * out_scale).  q/k/v int8 levels."""
    for t, nm in ((q, "q"), (k, "k"), (v, "v")):
        if t.dtype != torch.int8 or not t.is_cuda:
            raise S2FError(f"linear_attn: {nm} must be CUDA int8 levels")
    Cc = heads * d
    ws = torch.empty(int(_lib.lib().s2f_linear_attn_ws_bytes(n, heads, d)) // 4, dtype=torch.int32, device=q.device)
    out_ld = int(out_ld or Cc)
    out_s = torch.empty((n, Nq, out_ld), dtype=torch.int8, device=q.device)
    out_f = torch.empty((n, Nq, out_ld), dtype=torch.float32, device=q.device) if want_f32 else None
    e0 = _p0()
    check(_lib.lib().s2f_linear_attn(C.c_void_p(q.data_ptr()), C.c_void_p(k.data_ptr()), C.c_void_p(v.data_ptr()),
                                     _ptr(ws), _ptr(out_s), _ptr(out_f), n, Nq, Nk, heads, d, int(q_ld or Cc),
                                     int(kv_ld or Cc), out_ld, float(out_scale), float(d_max), _stream()), "s2f_linear_attn")
    _p1(e0, "linear_attn", 2.0 * n * heads * d * d * (Nq + Nk), n * (Nq + 2 * Nk) * Cc + _nb(out_s, out_f),
        f"{n} x Nq{Nq} Nk{Nk} h{heads} d{d}")
    return out_s, out_f


def dec_attn(q, k, v, *, n, Nq, Nk, heads, d, out_scale, mask=None, q_ld=None, kv_ld=None, out_ld=None, want_f32=False,
             d_max=D_MAX):
    """Masked spike attention in the reference's order: NI-LIF(((Q K^T) masked_fill(mask, 0)) V * out_scale).
    mask: bool / uint8 [n*heads, Nq, Nk] (True = masked, mmcv_spike/transformer.py:265-269, 349-353) or None."""
    for t, nm in ((q, "q"), (k, "k"), (v, "v")):
        if not isinstance(t, torch.Tensor) or t.dtype != torch.int8 or not t.is_cuda:
            raise S2FError(f"dec_attn: {nm} must be CUDA int8 levels (no CPU path)")
    Cc = heads * d
    mask_t = None                        # keeps the (possibly temporary) mask storage alive until the kernel is enqueued
    if mask is not None:
        if not mask.is_cuda or mask.numel() != n * heads * Nq * Nk:
            raise S2FError("dec_attn: mask must be a CUDA tensor of n*heads*Nq*Nk elements")
        mask_t = (mask.view(torch.uint8) if mask.dtype == torch.bool else mask).contiguous()
        mask = _ptr(mask_t, torch.uint8, "mask")
    out_ld = int(out_ld or Cc)
    out_s = torch.zeros((n, Nq, out_ld), dtype=torch.int8, device=q.device) if out_ld != Cc else \
        torch.empty((n, Nq, out_ld), dtype=torch.int8, device=q.device)
    out_f = torch.zeros((n, Nq, out_ld), dtype=torch.float32, device=q.device) if want_f32 else None
    e0 = _p0()
    check(_lib.lib().s2f_dec_attn(C.c_void_p(q.data_ptr()), C.c_void_p(k.data_ptr()), C.c_void_p(v.data_ptr()), mask,
                                  _ptr(out_s), _ptr(out_f), n, Nq, Nk, heads, d, int(q_ld or Cc), int(kv_ld or Cc), out_ld,
                                  float(out_scale), float(d_max), _stream()), "s2f_dec_attn")
    _p1(e0, "linear_attn", 4.0 * n * heads * d * Nq * Nk, n * (Nq + 2 * Nk) * Cc + _nb(out_s, out_f), f"{n} x Nq{Nq} Nk{Nk} masked")
    if mask_t is not None:
        mask_t.record_stream(torch.cuda.current_stream())
    return out_s, out_f


def dcnv3_gather(x, offset, mask, *, n, H, W, G, Cg, K=3, offset_scale=1.0, mask_scale=1.0 / NORM):
    """DCNv3 sampling core (dcnv3_core_pytorch, dcnv3_func.py:147-189)."""
    out = torch.empty((n, H, W, G * Cg), dtype=torch.float32, device=x.device)
    e0 = _p0()
    check(_lib.lib().s2f_dcnv3_gather(_ptr(x, torch.float32, "x"), _ptr(offset, torch.float32, "offset"),
                                      _ptr(mask, torch.int8, "mask"), float(mask_scale), _ptr(out), n, H, W, G, Cg, K,
                                      float(offset_scale), _stream()), "s2f_dcnv3_gather")
    _p1(e0, "dcn_gather", 0, _nb(x, offset, mask, out))
    return out


def upsample_add_lif(cur, prev, *, n, H, W, Hp, Wp, C_, want_f32=False, d_max=D_MAX):
    out_s = torch.empty((n, H, W, C_), dtype=torch.int8, device=cur.device)
    out_f = torch.empty((n, H, W, C_), dtype=torch.float32, device=cur.device) if want_f32 else None
    e0 = _p0()
    check(_lib.lib().s2f_upsample_add_lif(_ptr(cur, torch.float32, "cur"), _ptr(prev, torch.float32, "prev"),
                                          _ptr(out_s), _ptr(out_f), n, H, W, Hp, Wp, C_, float(d_max), _stream()),
          "s2f_upsample_add_lif")
    _p1(e0, "upsample_add_lif", 0, _nb(cur, prev, out_s, out_f), f"{n}x{H}x{W}x{C_}")
    return out_s, out_f


def sigmoid_lif(x, d_max=D_MAX):
    out = torch.empty(x.shape, dtype=torch.int8, device=x.device)
    e0 = _p0()
    check(_lib.lib().s2f_sigmoid_lif(_ptr(x, torch.float32, "x"), _ptr(out), x.numel(), float(d_max), _stream()),
          "s2f_sigmoid_lif")
    _p1(e0, "elementwise", 0, _nb(x, out))
    return out


def preprocess_u8(img, *, mean=None, std=None, swap_rb=False, size=None, pad_val=0.0, out=None):
    """SegDataPreProcessor arithmetic (data_preprocessor.py:121-126) + stack_batch padding for a uint8 batch
    [n,3,H,W] (CHW) or [n,H,W,3] (HWC).  Returns fp32 channels-last memory [n,Hp,Wp,3]."""
    if not isinstance(img, torch.Tensor) or not img.is_cuda or img.dtype != torch.uint8 or img.dim() != 4:
        raise S2FError("preprocess_u8: CUDA uint8 [n,3,H,W] or [n,H,W,3] tensor required (no CPU path)")
    if not img.is_contiguous():
        raise S2FError("preprocess_u8: img must be contiguous")
    chw = img.shape[1] == 3 and img.shape[3] != 3
    if not chw and img.shape[3] != 3:
        raise S2FError("preprocess_u8: three channels expected")
    n = int(img.shape[0])
    H, W = (int(img.shape[2]), int(img.shape[3])) if chw else (int(img.shape[1]), int(img.shape[2]))
    Hp, Wp = (max(H, int(size[0])), max(W, int(size[1]))) if size is not None else (H, W)
    if out is None:
        out = torch.empty((n, Hp, Wp, 3), dtype=torch.float32, device=img.device)
    elif tuple(out.shape) != (n, Hp, Wp, 3):
        raise S2FError("preprocess_u8: out must be fp32 [n,Hp,Wp,3]")
    m3 = s3 = None
    if mean is not None:
        # fp32 values of torch.tensor(mean) / torch.tensor(std) (data_preprocessor.py:88-91)
        m3 = (C.c_float * 3)(*[float(torch.tensor(float(v), dtype=torch.float32)) for v in mean])
        s3 = (C.c_float * 3)(*[float(torch.tensor(float(v), dtype=torch.float32)) for v in std])
    e0 = _p0()
    check(_lib.lib().s2f_preprocess_u8(C.c_void_p(img.data_ptr()), int(chw), _ptr(out, torch.float32, "out"), n, H, W, Hp, Wp,
                                       C.cast(m3, C.c_void_p) if m3 is not None else None,
                                       C.cast(s3, C.c_void_p) if s3 is not None else None, int(bool(swap_rb)),
                                       float(pad_val), _stream()), "s2f_preprocess_u8")
    _p1(e0, "elementwise", 0, _nb(img, out), f"preprocess {n}x{H}x{W}")
    return out


def stem_u8(img, w_packed, scale, shift_tab, *, Cout, want_f32=True, want_spike=True, d_max=D_MAX):
    """Preprocessor + 7x7/2 stem + BN (+ NI-LIF) from the uint8 batch ([n,3,H,W] or [n,H,W,3]) in one launch."""
    if not isinstance(img, torch.Tensor) or not img.is_cuda or img.dtype != torch.uint8 or img.dim() != 4 or not img.is_contiguous():
        raise S2FError("stem_u8: contiguous CUDA uint8 [n,3,H,W] or [n,H,W,3] batch required (no CPU path)")
    chw = img.shape[1] == 3 and img.shape[3] != 3
    n = int(img.shape[0])
    H, W = (int(img.shape[2]), int(img.shape[3])) if chw else (int(img.shape[1]), int(img.shape[2]))
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    of = torch.empty((n, Ho, Wo, Cout), dtype=torch.float32, device=img.device) if want_f32 else None
    os_ = torch.empty((n, Ho, Wo, Cout), dtype=torch.int8, device=img.device) if want_spike else None
    e0 = _p0()
    check(_lib.lib().s2f_stem_u8(C.c_void_p(img.data_ptr()), int(chw), _ptr(w_packed, torch.int8, "w_packed"),
                                 int(w_packed.shape[-1]), _ptr(scale, torch.float32, "scale"),
                                 _ptr(shift_tab, torch.float32, "shift_tab"), _ptr(of), _ptr(os_), n, H, W, int(Cout),
                                 float(d_max), _stream()), "s2f_stem_u8")
    _p1(e0, "stem_u8", 2.0 * n * Ho * Wo * Cout * 147, _nb(img, of, os_), f"{n}x{H}x{W} uint8 3->{Cout} k7s2")
    return of, os_


def peak_mma(kind="i8", iters=4096):
    """One launch of the tensor-pipe issue-rate kernel (csrc/peak.cu) on the current stream -> operations performed."""
    n = int(_lib.lib().s2f_peak_mma({"i8": 0, "bf16": 1}[kind], int(iters), _stream()))
    if n < 0:
        check(1, "s2f_peak_mma")
    return n


def level_hist(levels, hist=None):
    """hist[l] += number of elements at level l (l = 0..15); uint64-valued int64 tensor [16] on the device."""
    if levels.dtype != torch.int8 or not levels.is_cuda or not levels.is_contiguous():
        raise S2FError("level_hist: contiguous CUDA int8 levels required (no CPU path)")
    if hist is None:
        hist = torch.zeros(16, dtype=torch.int64, device=levels.device)
    e0 = _p0()
    check(_lib.lib().s2f_level_hist(C.c_void_p(levels.data_ptr()), int(levels.numel()), _ptr(hist, torch.int64, "hist"),
                                    _stream()), "s2f_level_hist")
    _p1(e0, "elementwise", 0, _nb(levels))
    return hist


def semantic_tail(mask_pred, cls, *, n, Q, K, h, w, H, W, want_logits=True, want_labels=False):
    """Tensor-core tail: softmax(cls)[..., :-1] x sigmoid(bilinear(mask_pred)) -> (logits [n,K,H,W] | None,
    labels uint8 [n,H,W] | None).  Falls outside the tcgen05 kernel's shape range only for Q > 128 or K > 256."""
    if Q > 128 or K > 256:
        if want_labels:
            raise S2FError("semantic_tail: fused labels need Q <= 128 and K <= 256")
        return semantic_tail_simt(mask_pred, cls, n=n, Q=Q, K=K, h=h, w=w, H=H, W=W), None
    lib = _lib.lib()
    logits = torch.empty((n, K, H, W), dtype=torch.float32, device=cls.device) if want_logits else None
    labels = torch.empty((n, H, W), dtype=torch.uint8, device=cls.device) if want_labels else None
    ws = torch.empty(lib.s2f_semantic_tail_ws_bytes(n, K), dtype=torch.uint8, device=cls.device)
    e0 = _p0()
    check(lib.s2f_semantic_tail_tc(_ptr(mask_pred, torch.float32, "mask_pred"), _ptr(cls, torch.float32, "cls"),
                                   _ptr(logits), _ptr(labels), _ptr(ws), n, Q, K, h, w, H, W, _stream()),
          "s2f_semantic_tail_tc")
    _p1(e0, "semantic_tail", 2.0 * n * H * W * Q * K, _nb(mask_pred, logits, labels))
    return logits, labels


def semantic_tail_simt(mask_pred, cls, *, n, Q, K, h, w, H, W):
    """CUDA-core version (any Q / K): softmax(cls)[..., :-1] x sigmoid(bilinear(mask_pred)) -> logits [n, K, H, W]."""
    logits = torch.empty((n, K, H, W), dtype=torch.float32, device=cls.device)
    prob = torch.empty((n, Q, K), dtype=torch.float32, device=cls.device)
    e0 = _p0()
    check(_lib.lib().s2f_semantic_tail(_ptr(mask_pred, torch.float32, "mask_pred"), _ptr(cls, torch.float32, "cls"),
                                       _ptr(logits), _ptr(prob), n, Q, K, h, w, H, W, _stream()), "s2f_semantic_tail")
    _p1(e0, "semantic_tail", 2.0 * n * H * W * Q * K, _nb(mask_pred, logits))
    return logits
