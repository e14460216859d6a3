"""Teacher-forced parity probe (TEST INFRASTRUCTURE ONLY).

The random-init network is chaotic: a single flipped spike grows ~5-10x per layer, so the reference
disagrees with *itself* end to end when run in fp64 instead of fp32 (tests/test_chaos.py).
Parity is therefore established unit by unit: the oracle port is run once, every neuron's
pre-activation and levels (and a few marked real tensors) are recorded, and the CUDA engine is run
with this probe, which at every neuron

  * compares the engine's levels with the oracle's (flips are only tolerated where the oracle's own
    pre-activation sits within `tie_tol` of a rounding boundary k + 0.5),
  * compares real tensors within `rtol` of the tensor's scale,
  * and then REPLACES the engine's tensor by the oracle's, so each fused kernel is driven by exactly
    the inputs the reference saw ("bit-exact given identical pre-activations").
"""
from __future__ import annotations

import torch


def record_oracle(P, cfg, img):
    """Run the port once; returns dict name -> (pre fp32, levels int8) and dict of marked real tensors, logits."""
    from . import port

    taps = {}

    def tap(name, pre, s):
        taps.setdefault(name, []).append((pre.detach(), s.to(torch.int8)))

    cx = port.Ctx(P, tap=tap)
    cx.marks = {}
    with torch.no_grad():
        logits = port.predict(cx, cfg, img)
    # every neuron is called once per forward in this model
    return {k: v[0] for k, v in taps.items()}, cx.marks, logits


def to_engine_layout(o: torch.Tensor, mine_shape, layout: str) -> torch.Tensor:
    """Oracle tensor -> the engine's layout (see engine.NullProbe).  The engine may carry zero-padded channels
    (last dim rounded up to 16 for TMA); the oracle tensor is zero-padded to match."""
    mine_shape = tuple(mine_shape)
    if isinstance(layout, tuple):                       # ("cm_heads", heads, d, dp): per-head channel padding
        _, heads, d, dp = layout
        core = to_engine_layout(o, mine_shape[:-1] + (heads * d,), "cm")
        out = torch.zeros(mine_shape[:-1] + (heads, dp), dtype=core.dtype)
        out[..., :d] = core.reshape(mine_shape[:-1] + (heads, d))
        return out.reshape(mine_shape)
    c_ref = o.shape[1] if layout == "cm" else o.shape[-1]
    if layout in ("cm", "same") and mine_shape[-1] > c_ref and (mine_shape[-1] - c_ref) < 16 and \
            o.numel() // c_ref * mine_shape[-1] == int(torch.tensor(mine_shape).prod()):
        core = to_engine_layout(o, mine_shape[:-1] + (c_ref,), layout)
        return torch.nn.functional.pad(core, (0, mine_shape[-1] - c_ref))
    if layout == "same":
        return o.reshape(mine_shape)
    if layout == "cm":
        n, c = o.shape[0], o.shape[1]
        return o.reshape(n, c, -1).permute(0, 2, 1).reshape(mine_shape)
    if layout == "reint_T":      # oracle [n, nq, C] read as [n, C, nq]; engine holds its transpose [n, nq, C]
        n, nq, c = o.shape
        return o.reshape(n, c, nq).permute(0, 2, 1).reshape(mine_shape)
    raise ValueError(layout)


class TeacherProbe:
    active = True

    # tie_tol: the reference accumulates its convolutions in fp32 (K up to 3312 terms, then a BatchNorm scale of
    # ~10), so ITS pre-activations carry ~1e-5 of rounding noise; a flip within 1e-4 of a rounding boundary is
    # indistinguishable from that noise.  The engine's int32 accumulation is exact (error 2^-21 per weight row).
    def __init__(self, taps, marks, device, prefix="", force=True, tie_tol=1e-4, rtol=2e-5, log=None):
        self.taps, self.marks, self.device, self.prefix = taps, marks, device, prefix
        self.force, self.tie_tol, self.rtol = force, tie_tol, rtol
        self.log = log if log is not None else []       # shared between scoped copies

    def scoped(self, prefix):
        return TeacherProbe(self.taps, self.marks, self.device, self.prefix + prefix, self.force, self.tie_tol,
                            self.rtol, self.log)

    # -- helpers
    def _entry(self, kind, name, **kw):
        e = dict(kind=kind, name=name, **kw)
        self.log.append(e)
        return e

    def spike(self, name, t, layout="cm"):
        full = self.prefix + name
        if full not in self.taps:
            self._entry("spike", full, status="unknown")
            return t
        pre, lv = self.taps[full]
        want = to_engine_layout(lv, t.shape, layout).contiguous()
        pre_m = to_engine_layout(pre, t.shape, layout)
        got = t.detach().cpu()
        diff = got != want
        nflip = int(diff.sum())
        # a flip is explainable iff the oracle's own pre-activation is within tie_tol of a rounding boundary
        frac = pre_m - torch.floor(pre_m)
        near = ((frac - 0.5).abs() <= self.tie_tol * pre_m.abs().clamp(min=1.0)) & (pre_m > -self.tie_tol) & \
               (pre_m < 8.0 + 0.5 + self.tie_tol)
        unexplained = int((diff & ~near).sum())
        maxdev = int((got.int() - want.int()).abs().max()) if got.numel() else 0
        # distance of the worst flipped element from its rounding boundary, relative to max(1, |pre|)
        gap = float(((frac - 0.5).abs() / pre_m.abs().clamp(min=1.0))[diff].max()) if nflip else 0.0
        n_ref = lv.numel()                                  # reference element count (without channel padding)
        self._entry("spike", full, numel=n_ref, flips=nflip, unexplained=unexplained, maxdev=maxdev,
                    near_ties=int(near.sum()), rate=float(want.float().mean()), worst_gap=gap)
        return want.to(self.device) if self.force else t

    def real(self, name, t, layout="cm"):
        full = self.prefix + name
        if full in self.taps:
            ref = self.taps[full][0]
        elif full in self.marks:
            ref = self.marks[full]
        else:
            self._entry("real", full, status="unknown")
            return t
        want = to_engine_layout(ref, t.shape, layout).contiguous()
        got = t.detach().cpu()
        scale = float(want.abs().max().clamp(min=1e-6))
        err = float((got - want).abs().max())
        self._entry("real", full, numel=got.numel(), max_abs_err=err, scale=scale, rel=err / scale)
        return want.to(self.device) if self.force else t

    # -- summaries
    def summary(self):
        sp = [e for e in self.log if e["kind"] == "spike" and "flips" in e]
        re_ = [e for e in self.log if e["kind"] == "real" and "rel" in e]
        unk = [e["name"] for e in self.log if e.get("status") == "unknown"]
        return dict(
            neurons=len(sp), spike_elems=sum(e["numel"] for e in sp), flips=sum(e["flips"] for e in sp),
            unexplained=sum(e["unexplained"] for e in sp), maxdev=max([e["maxdev"] for e in sp] or [0]),
            reals=len(re_), worst_rel=max([e["rel"] for e in re_] or [0.0]),
            worst_real=max(re_, key=lambda e: e["rel"])["name"] if re_ else None, unknown=unk)


class ObserverProbe:
    """Free-running parity: taps every tensor the PRODUCTION path hands over (engine.NullProbe docstring) without
    replacing anything or switching code paths; under CUDA-graph capture a tap is one extra copy node.
    `compare(taps)` then counts, neuron by neuron, the levels that differ from the oracle's."""
    active = False
    observe = True

    def __init__(self, prefix="", spikes=None, reals=None):
        self.prefix = prefix
        self.spikes = {} if spikes is None else spikes
        self.reals = {} if reals is None else reals

    def scoped(self, prefix):
        return ObserverProbe(self.prefix + prefix, self.spikes, self.reals)

    def spike(self, name, t, layout="cm"):
        self.spikes[self.prefix + name] = (t.detach().clone(memory_format=torch.contiguous_format), layout)
        return t

    def real(self, name, t, layout="cm"):
        self.reals[self.prefix + name] = (t.detach().clone(memory_format=torch.contiguous_format), layout)
        return t

    def compare(self, taps, marks=None):
        """-> dict(neurons, spike_elems, flips, maxdev, per_neuron=[(name, flips, numel)] in the engine's call order,
        missing=[oracle neurons the engine never showed], reals=[(name, rel err)])."""
        per, total, flips, maxdev = [], 0, 0, 0
        for name, (t, layout) in self.spikes.items():
            if name not in taps:
                continue
            lv = taps[name][1]
            got = t.cpu()
            if layout == "same" and got.dim() == lv.dim() and got.shape[0] == 1 and lv.shape[0] > 1:
                lv = lv[-1:]                                  # last-only SDME: the engine keeps decoder output [-1]
            want = to_engine_layout(lv, got.shape, layout)
            diff = got != want
            f = int(diff.sum())
            if f:
                maxdev = max(maxdev, int((got.int() - want.int()).abs().max()))
            per.append((name, f, lv.numel()))
            total += lv.numel(); flips += f
        reals = []
        for name, (t, layout) in self.reals.items():
            ref = taps[name][0] if name in taps else (marks or {}).get(name)
            if ref is None:
                continue
            got = t.cpu()
            if layout == "same" and got.dim() == ref.dim() and got.shape[0] == 1 and ref.shape[0] > 1:
                ref = ref[-1:]
            want = to_engine_layout(ref, got.shape, layout)
            reals.append((name, float((got - want).abs().max() / want.abs().max().clamp(min=1e-6))))
        seen = {n for n, _, _ in per}
        return dict(neurons=len(per), spike_elems=total, flips=flips, maxdev=maxdev, per_neuron=per,
                    missing=[n for n in taps if n not in seen], reals=reals)
