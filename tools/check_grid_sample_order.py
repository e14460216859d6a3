"""Build-container check: which fp32 operation order does torch's CPU `F.grid_sample` (the oracle's DCNv3 sampler,
dcnv3_func.py:179-180) use?  Emulates the candidates in numpy and compares bit for bit on random points.
Result (torch 2.11 CPU): ix = fma(g + 1, W / 2, -0.5); value = fma(se, v_se, fma(sw, v_sw, fma(ne, v_ne, nw * v_nw)))
-- the order `dcnv3_kernel` (csrc/spatial.cu) follows.

    python tools/check_grid_sample_order.py
"""
import numpy as np
import torch
import torch.nn.functional as F

f, d = np.float32, np.float64


def fl(x):
    return x.astype(f)


def fma(a, b, c):
    return (np.asarray(a).astype(d) * np.asarray(b).astype(d) + np.asarray(c).astype(d)).astype(f)


def main(W=34, H=34, N=400000):
    torch.manual_seed(0)
    v = torch.rand(1, 1, H, W)
    gx, gy = (torch.rand(N) * 2 - 1) * 0.95, (torch.rand(N) * 2 - 1) * 0.95
    out = F.grid_sample(v, torch.stack([gx, gy], -1).view(1, 1, N, 2), mode="bilinear", padding_mode="zeros",
                        align_corners=False).view(-1).numpy()
    img = v[0, 0].numpy()
    g1, g2 = gx.numpy().astype(f), gy.numpy().astype(f)

    def val(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        return np.where(ok, img[np.clip(yy, 0, H - 1).astype(int), np.clip(xx, 0, W - 1).astype(int)], 0).astype(f)

    def sample(ix, iy):
        x0, y0 = np.floor(ix), np.floor(iy)
        w, n = fl(ix - x0), fl(iy - y0)
        e, s = fl(f(1) - w), fl(f(1) - n)
        r = fl(val(y0, x0) * fl(s * e))
        for a, b in ((val(y0, x0 + 1), fl(s * w)), (val(y0 + 1, x0), fl(n * e)), (val(y0 + 1, x0 + 1), fl(n * w))):
            r = fma(a, b, r)
        return r

    sep = sample(fl(fl(fl(g1 + f(1)) * f(W / 2)) - f(0.5)), fl(fl(fl(g2 + f(1)) * f(H / 2)) - f(0.5)))
    one = sample(fma(fl(g1 + f(1)), f(W / 2), f(-0.5)), fma(fl(g2 + f(1)), f(H / 2), f(-0.5)))
    print(f"product rounded separately: {float((sep == out).mean()):.4f} bit-equal; fused multiply-add: {float((one == out).mean()):.4f}")
    assert float((one == out).mean()) == 1.0


if __name__ == "__main__":
    main()
