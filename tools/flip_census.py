"""GPU tool: where do spike flips between the CUDA path and the oracle come from, and how do they grow?

    python tools/flip_census.py [ade20k_stable|ade20k|tiny_stable] [HxW] [batch]  -> gpurun_out/flip_census_<name>.json

seeds[n]  : flips of neuron n when it is fed the ORACLE's inputs (teacher forcing) -- arithmetic differences only;
free[n]   : flips of neuron n in the free-running production path (observer probe) -- seeds + everything they caused.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import spike2former_b200 as s2f  # noqa: E402
from oracle import probe, weights  # noqa: E402
from spike2former_b200 import engine, synth  # noqa: E402


def main(name="ade20k_stable", hw="512x512", batch=1):
    h, w = (int(v) for v in hw.split("x"))
    cfg = getattr(s2f.configs, name.split("_")[0])()
    P = synth.synthetic_checkpoint(name, cfg)
    img = weights.test_image(cfg, h, w, batch=int(batch))
    taps, marks, ref = probe.record_oracle(P, cfg, img)
    seg = s2f.build_segmentor(cfg)
    seg.load_state_dict(P, strict=True)
    seg = seg.cuda()
    tp = probe.TeacherProbe(taps, marks, torch.device("cuda"))
    with torch.no_grad():
        engine.segmentor_logits(seg, img.cuda(), tp)
    seeds = {e["name"]: (e["flips"], e["numel"], e["worst_gap"]) for e in tp.log if e["kind"] == "spike" and "flips" in e}
    reals = {e["name"]: e["rel"] for e in tp.log if e["kind"] == "real" and "rel" in e}
    obs = probe.ObserverProbe()
    with torch.no_grad():
        out = engine.segmentor_logits(seg, img.cuda(), obs).cpu()
    r = obs.compare(taps, marks)
    free = {n: f for n, f, _ in r["per_neuron"]}
    agree = float((out.argmax(1) == ref.argmax(1)).float().mean())
    rows = [(n, seeds.get(n, (None,))[0], free.get(n), seeds.get(n, (0, 0))[1]) for n in taps if n in free or n in seeds]
    print(f"{name} {hw} batch {batch}: seeds {sum(v[0] for v in seeds.values())}, free-running flips {r['flips']} of {r['spike_elems']}, "
          f"argmax agreement {agree:.5f}, logits rel L2 {float((out - ref).norm() / ref.norm()):.3e}")
    for n, s, f, m in rows:
        if (s or 0) > 0 or (f or 0) > 0:
            print(f"  {n:<75s} seeds {s!s:>7} free {f!s:>8} of {m}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dict(name=name, hw=hw, batch=batch, rows=rows, agree=agree, teacher_reals=reals),
              open(os.path.join(ROOT, "gpurun_out", f"flip_census_{name}.json"), "w"))


if __name__ == "__main__":
    main(*sys.argv[1:])
