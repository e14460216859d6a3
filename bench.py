#!/usr/bin/env python
"""bench.py -- Spike2Former 512x512 batch-sharded inference on N B200s (BASELINE.json metric: imgs/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

One "step" = one forward pass of the whole hot path (SDTv2 backbone + DCN pixel decoder + MaskFormer
head -> seg logits [B,150,512,512]) over one batch of B synthetic 512x512 images per GPU.
  value : whole-job images/s with the input batches already resident in HBM (CUDA events, max over ranks)
  e2e   : the same through the public API with HOST buffers: pinned uint8 images (the dataset format the reference's
          SegDataPreProcessor receives) -> H2D -> normalise + forward -> argmax label map (uint8) -> D2H, every step,
          inside the timed region
  roofline : the dominant kernel class (spike GEMM / conv) timed live with CUDA events on the launch stream, on the
          reference layers' ALGORITHMIC FLOPs (executed FLOPs beside it), plus a per-class table
  tensor_peaks : tcgen05 issue-rate ceilings of kind::i8 and kind::f16(bf16) measured in this run (csrc/peak.cu)
  kernels : BASELINE.json config 2 micro-benchmarks (NI-LIF D=8 / D=4, spike GEMM, spike-driven attention C=512 d=64)
  batches : the same model at 1 and 8 images per GPU (the reference's own timing protocol is batch 1)
  cityscapes_1024x2048 : BASELINE.json config 4 with its own value / e2e / per-class roofline
  cpu_baseline : the oracle port (CPU restatement of the reference, pinned bit-exact to it) on this box's host
          cores over a bounded sample, rank 0, N=1 only
`--impl reference` times that CPU path alone with the same JSON contract.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

H = W = 512
WORKLOAD = "Spike2Former SDTv2 + DCN pixel decoder, ADE20K-shape 512x512 inference (150 classes, 100 queries)"
ALG_GFLOP_PER_IMG = 131.0      # 142 GFLOP minus the six discarded mask einsums (inference uses [-1])
CITY_H, CITY_W = 1024, 2048    # BASELINE.json config 4: Cityscapes shape (19 classes)
NBUF = 4                       # rotating input batches


def workload_config(B, world):
    """The `config` object of the JSON line -- the same for the CUDA arm and the reference arm."""
    return {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": world * B, "parallelism": f"dp{world}",
            "weights": "seeded random init + shipped calibration statistics (spike2former_b200/synth.py)",
            "l2": f"{NBUF} rotating input batches; per-step working set >> 126 MB L2",
            "value_input": "uint8 [B,3,512,512] resident in HBM (SegDataPreProcessor + encode_decode -> fp32 logits [B,150,512,512])",
            "e2e_input": "uint8 [B,3,512,512] pinned host memory -> SegDataPreProcessor + predict -> uint8 labels -> host",
            "alg_gflop_per_image": ALG_GFLOP_PER_IMG}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return dict(sm_mhz=statistics.median(self.samples) if self.samples else None, sm_max_mhz=self.max_mhz,
                    reasons=sorted(self.reasons), samples=len(self.samples))


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_run(steps, warmup):
    """The reference's CPU implementation of the path: oracle/port.py (bit-exact restatement, see tests/golden).
    One step = preprocess + forward + argmax of ONE 512x512 image (a bounded sample of the batched workload)."""
    from oracle import port
    from spike2former_b200 import configs, synth

    cfg = configs.ade20k()
    P = synth.synthetic_checkpoint("ade20k", cfg)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    u8 = [torch.randint(0, 256, (3, H, W), generator=g, dtype=torch.uint8)]
    dp = cfg["data_preprocessor"]

    def one():       # SegDataPreProcessor + encode_decode + argmax, as BaseSegmentor.test_step does
        x = port.data_preprocess(u8, mean=dp["mean"], std=dp["std"], bgr_to_rgb=dp["bgr_to_rgb"])
        return port.predict(port.Ctx(P), cfg, x).argmax(1)

    with torch.no_grad():
        for _ in range(warmup):
            one()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        dt = time.perf_counter() - t0
    return dict(value=steps / dt, unit="images/s", cores=cores, kind="port",
                sample=f"{steps} x (preprocess + forward + argmax) of one 512x512 uint8 image (batch 1), fp32, torch CPU {torch.__version__}, "
                       f"{cores} threads, after {warmup} warm-up"), dt / steps * 1e3


# ----------------------------------------------------------------------------------------------- micro-benchmarks
class KernelTimer:
    def __init__(self, dev):
        self.flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)

    def __call__(self, fn, iters=10):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            self.flush.view(torch.int64).sum()             # evict the working set, leave clean lines in L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(400000)                      # the launch is queued before the GPU reaches e0
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts) * 1e-3


def tensor_peaks(dev):
    """Tensor-pipe ceilings measured in this run, the way MEASURED_PEAKS.json measures bf16 (best of 10 = burst, back to
    back for ~3 s = sustained): (i) tcgen05 issue rate of kind::i8 and kind::f16 (csrc/peak.cu: operands resident in
    shared memory, nothing but the tensor pipe can limit it), (ii) cuBLASLt int8 GEMM 8192^3 through torch._int_mm."""
    from spike2former_b200 import ops

    out = {}
    for kind in ("i8", "bf16"):
        iters = 4096 if kind == "i8" else 8192               # ~5e12 ops per launch (five 8192^3 GEMMs' worth)
        for _ in range(3):
            n_ops = ops.peak_mma(kind, iters)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); n_ops = ops.peak_mma(kind, iters); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(10, int(3000.0 / best))
        e0.record()
        for _ in range(reps):
            ops.peak_mma(kind, iters)
        e1.record()
        torch.cuda.synchronize()
        out[f"tcgen05_{kind}_issue_rate"] = dict(burst_tops=n_ops / best / 1e9, sustained_tops=n_ops * reps / e0.elapsed_time(e1) / 1e9,
                                                 ops_per_launch=n_ops)
    try:
        a = torch.randint(-8, 8, (8192, 8192), dtype=torch.int8, device=dev)
        b = torch.randint(-8, 8, (8192, 8192), dtype=torch.int8, device=dev).t()
        for _ in range(3):
            torch._int_mm(a, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch._int_mm(a, b); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out["cublaslt_int8_8192^3"] = dict(burst_tops=2.0 * 8192 ** 3 / best / 1e9)
    except Exception as e:  # pragma: no cover
        out["cublaslt_int8_8192^3"] = dict(unavailable=str(e)[:120])
    return out


def kernel_microbench(dev, pk, tp):
    """BASELINE.json config 2 on one GPU (B=64, N=1024 tokens, C=512, 8 heads, d=64, T=1): fused NI-LIF (D=8 and D=4),
    the MLP spike GEMM 512->2048 and the spike-driven attention Q(K^T V)."""
    from spike2former_b200 import ops

    timed = KernelTimer(dev)
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(64, 1024, 512, generator=g) * 12 - 2).to(dev)
    lv = torch.empty(x.shape, dtype=torch.int8, device=dev)
    nb = x.numel() * 5

    def lif_entry(t, label):
        return dict(workload=label, us=t * 1e6, gbs=nb / t / 1e9, algorithmic_bytes=nb, frac_of_measured_hbm=nb / t / 1e9 / pk["hbm"],
                    frac_of_8tbs_nominal=nb / t / 1e9 / 8000.0)

    nilif = lif_entry(timed(lambda: ops.nilif(x, out=lv)), "NI-LIF B=64 N=1024 C=512 T=1 D=8 (Q_IFNode / Quant, /8)")
    bn_s, bn_b = (torch.rand(512, generator=g) + 0.5).to(dev), torch.randn(512, generator=g).to(dev)
    t = timed(lambda: ops.nilif(x, scale=bn_s, shift=bn_b, out=lv))
    nilif["with_folded_bn_affine"] = dict(us=t * 1e6, gbs=nb / t / 1e9, frac_of_measured_hbm=nb / t / 1e9 / pk["hbm"])
    nilif_d4 = lif_entry(timed(lambda: ops.nilif(x, out=lv, d_max=4.0, norm=4.0)),
                         "NI-LIF B=64 N=1024 C=512 T=1 D=4 (Quant4 / Multispike_norm, /4): BASELINE config 2 as written")
    n, Hh, Ww, cin, cout = 64, 32, 32, 512, 2048
    a = torch.randint(0, 9, (n, Hh, Ww, cin), generator=g, dtype=torch.int8).to(dev)
    w = torch.randn(cout, cin, generator=g) / cin ** 0.5
    packed, rowscale = ops.pack_weights_i8(w, 1, cin, 3)
    packed, sc, sh = packed.to(dev), (rowscale / 8).to(dev), torch.zeros(cout, device=dev)
    t = timed(lambda: ops.gemm_tc(a, packed, n=n, H=Hh, W=Ww, Cin=cin, Cout=cout, scale=sc, shift=sh, pieces=3,
                                  want_spike=True))
    alg = 2.0 * n * Hh * Ww * cin * cout / t / 1e12
    i8 = tp["tcgen05_i8_issue_rate"]["burst_tops"]
    gemm = dict(workload="spike GEMM fc1 512->2048, 65536 tokens, int8 spikes x 3 int8 weight planes, NI-LIF epilogue",
                us=t * 1e6, tflops_algorithmic=alg, algorithmic_vs_measured_bf16_burst=alg / pk["bf16"],
                int8_tops_executed=3.0 * alg, executed_vs_measured_i8_issue_rate=3.0 * alg / i8)
    # spike-driven attention of one MS_Block at config 2: q, k, v int8 levels [64, 1024, 512], 8 heads x 64
    heads, d, N = 8, 64, 1024
    q, k, v = (torch.randint(0, 3, (64, N, heads * d), generator=g, dtype=torch.int8).to(dev) for _ in range(3))
    t = timed(lambda: ops.linear_attn(q, k, v, n=64, Nq=N, Nk=N, heads=heads, d=d, out_scale=d ** -0.5 / 512))
    fl = 2.0 * 64 * heads * d * d * 2 * N                    # K^T V and Q (K^T V): SURVEY.md section 8d, 8.6 GFLOP
    byts = 4.0 * 64 * N * heads * d                          # q, k, v read + levels written, 1 B each
    sdsa = dict(workload="SDSA Q(K^T V) + NI-LIF, B=64 N=1024 C=512 8 heads d=64 (sdtv2.py:335-336)", us=t * 1e6,
                tflops=fl / t / 1e12, gbs=byts / t / 1e9, frac_of_measured_hbm=byts / t / 1e9 / pk["hbm"],
                bound="hbm: 268 MB of int8 operands against 8.6 GFLOP (32 FLOP/B)")
    return dict(nilif_cfg2=nilif, nilif_cfg2_d4=nilif_d4, gemm_cfg2=gemm, sdsa_cfg2=sdsa)


# ----------------------------------------------------------------------------------------------- model timing
class Harness:
    """value / e2e timing of one segmentor at one (batch, shape), on this rank."""

    def __init__(self, seg, B, h, w, dev, dist, seed):
        self.seg, self.B, self.h, self.w, self.dev, self.dist = seg, B, h, w, dev, dist
        g = torch.Generator().manual_seed(seed)
        self.host = [torch.randint(0, 256, (B, 3, h, w), generator=g, dtype=torch.uint8).pin_memory() for _ in range(NBUF)]
        # `value`: the same uint8 images (the format the reference's SegDataPreProcessor receives from the dataset) already
        # resident in HBM; the preprocessor is folded into the tensor-core stem, so no fp32 image is ever materialised
        self.devin = [x.to(dev) for x in self.host]
        # end-to-end: pinned uint8 host batch -> H2D (copy stream, double buffered so the copy of step i+1 overlaps the
        # forward of step i) -> fused preprocessor + forward with fused argmax (one CUDA graph) -> D2H of the uint8 label
        # map; all of it inside the timed region
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.stage = [torch.empty(B, 3, h, w, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.ev_ready = [torch.cuda.Event() for _ in range(2)]
        self.ev_free = [torch.cuda.Event() for _ in range(2)]
        self.host_outs = [torch.empty(B, h, w, dtype=torch.uint8).pin_memory() for _ in range(2)]

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def step_resident(self, i, labels=False):
        with torch.no_grad():
            x = self.devin[i % NBUF]
            return self.seg.predict_labels(x) if labels else self.seg.encode_decode(x)

    def step_e2e(self, i):
        j = i & 1
        main = torch.cuda.current_stream()
        with torch.no_grad():
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self.ev_free[j])
                self.stage[j].copy_(self.host[i % NBUF], non_blocking=True)
                self.ev_ready[j].record(self.copy_stream)
            main.wait_event(self.ev_ready[j])
            labels = self.seg.predict_labels(self.stage[j])
            self.ev_free[j].record(main)
            self.host_outs[j].copy_(labels, non_blocking=True)
        return self.host_outs[j]

    def timed(self, fn, steps, warmup):
        """W warm-up steps, then exactly K steps between barriers; CUDA events; max over ranks -> ms for the K steps."""
        for i in range(warmup):
            fn(i)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        self.barrier()
        from spike2former_b200 import dist as s2f_dist

        return s2f_dist.max_over_ranks(e0.elapsed_time(e1), self.dev)


def build(name, dev):
    import spike2former_b200 as s2f
    from spike2former_b200 import synth

    cfg = getattr(s2f.configs, name)()
    seg = s2f.build_segmentor(cfg)
    seg.load_state_dict(synth.synthetic_checkpoint(name, cfg), strict=True)
    seg = seg.to(dev)
    seg.alias_graph_output = True        # serving loop: the result is consumed (copied out / dropped) before the next call,
    return seg                           # so the graph's static output buffer is handed out without a 5 GB copy per step


def training_step_bench(dev, world, rank, dist, B, precision, steps=5, warmup=2, graph=True):
    """BASELINE.json config 5: one surrogate-gradient training step of the ADE20K model (batch 6 per GPU as in the
    reference config :181-182, 512x512 crops, synthetic labels), forward + 21 losses + backward + gradient all-reduce over
    NCCL (N > 1) + clip + AdamW, timed with CUDA events between barriers, max over ranks."""
    import spike2former_b200 as s2f
    from spike2former_b200 import dist as s2f_dist, synth, train

    cfg = s2f.configs.ade20k()
    seg = s2f.build_segmentor(cfg)
    seg.load_state_dict(synth.synthetic_checkpoint("ade20k", cfg), strict=True)
    seg = seg.to(dev)
    step = train.TrainStep(seg, precision=precision, graph=graph)
    g = torch.Generator().manual_seed(500 + rank)
    imgs = [torch.randn(B, 3, H, W, generator=g).to(dev) for _ in range(2)]
    gts = [torch.randint(0, 150, (B, 1, H // 16, W // 16), generator=g).repeat_interleave(16, 2).repeat_interleave(16, 3).to(dev)
           for _ in range(2)]
    for i in range(warmup):
        step(imgs[i & 1], gts[i & 1])
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    timing = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        losses = step(imgs[i & 1], gts[i & 1], timing=timing)
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = s2f_dist.max_over_ranks(e0.elapsed_time(e1), dev) / steps
    out = dict(workload="Spike2Former ADE20K training step (surrogate-gradient NI-LIF backward, 21 loss terms, AdamW, clip 0.01)",
               batch_per_gpu=B, n_gpus=world, precision=precision, cuda_graph_forward_backward=graph, steps=steps, ms_per_step=ms, images_per_second=world * B / (ms / 1e3),
               phases_ms={k: v / steps for k, v in timing.items()}, total_loss=float(sum(losses.values())),
               params=sum(p.numel() for p in seg.parameters()),
               allreduce_bytes_per_step=step.buckets.bytes if step.buckets is not None else 0,
               allreduce_buckets=len(step.buckets.buckets) if step.buckets is not None else 0,
               host_syncs_per_step="1 D2H of all 7 x B cost matrices + 1 H2D of the assignments (reference: 7 x B)",
               peak_memory_gb=torch.cuda.max_memory_allocated(dev) / 2 ** 30)
    del step, seg
    torch.cuda.empty_cache()
    return out


def class_rooflines(roof, pk, tp):
    """Per kernel class: share of the step and the fraction of the roof that bounds it."""
    from spike2former_b200 import engine

    out = {}
    for k, c in roof["classes"].items():
        e = dict(ms=c["ms"], launches=c["launches"])
        if k.startswith("gemm") or k in ("semantic_tail", "stem_u8"):
            e.update(tflops_algorithmic=c["tflops_algorithmic"], tflops_executed=c["tflops_executed"],
                     frac_of_bf16_sustained=round(c["tflops_algorithmic"] / pk["bf16_sustained"], 3))
            if k == "gemm_tc":
                e["int8_executed_frac_of_i8_issue_rate"] = round(3 * c["tflops_executed"] / tp["tcgen05_i8_issue_rate"]["sustained_tops"], 3) if tp else None
        if k in engine.HBM_CLASSES or k in ("semantic_tail", "stem_u8"):
            e.update(gbs=c["gbs"], frac_of_hbm=round(c["gbs"] / pk["hbm"], 3))
        out[k] = e
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step (throughput saturates at 64: 32 -> 2223, 64 -> 2324, 128 -> 2334 images/s)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip micro-benchmarks, peaks, batch 1 / 8 and the CPU baseline")
    ap.add_argument("--city-batch", type=int, default=8, help="images per GPU for the 1024x2048 config (0 = skip)")
    ap.add_argument("--train-batch", type=int, default=6, help="images per GPU of the training-step measurement (config 5; 0 = skip)")
    ap.add_argument("--train-precision", default="tf32", choices=["fp32", "tf32", "bf16"])
    args = ap.parse_args()
    from spike2former_b200 import dist as s2f_dist

    rank, world, local = s2f_dist.env_world()

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warm = max(1, args.steps), max(0, args.warmup)
        cb, ms = cpu_reference_run(steps, warm)
        print(json.dumps({
            "impl": "reference", "metric": "images_per_second", "value": cb["value"], "unit": "images/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.batch, max(1, args.gpus)),
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    from spike2former_b200 import engine, ops

    warm = max(args.warmup, 3)                 # >= 3 warm-ups: the second sighting of a shape captures its CUDA graph
    seg = build("ade20k", dev)
    B = args.batch
    hs = Harness(seg, B, H, W, dev, dist, 1000 + rank)

    # ------------------------------------------------------------------ device-resident timing (the headline `value`)
    for i in range(warm):
        hs.step_resident(i)
    hs.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ms_max = hs.timed(hs.step_resident, args.steps, 0)
    clocks = sampler.stop()
    launches = args.steps * sum(g.launches for k, g in seg._graphs.items() if not k[2])
    value = world * B * args.steps / (ms_max / 1e3)

    # ------------------------------------------------------------------ end-to-end (host buffers) timing
    e2e_ms = hs.timed(hs.step_e2e, args.steps, 4)
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)

    extras = rank == 0 and world == 1 and not args.no_extras
    pk = peaks()
    # ------------------------------------------------------------------ roofline of the dominant kernel class
    roof = engine.profile_dominant(seg, hs.devin[0], steps=min(args.steps, 3)) if rank == 0 else None
    tp = tensor_peaks(dev) if extras else None
    micro = kernel_microbench(dev, pk, tp) if extras else None

    batches = None
    if extras:        # the reference's own protocol is batch 1 (tools/analysis_tools/benchmark.py:57-110); SURVEY 8d asks for 1 / 8 / 32; the headline runs 64
        batches = {}
        for b in (1, 8, 32):
            seg._graphs.clear()
            h2 = Harness(seg, b, H, W, dev, None, 2000 + b)
            ms_b = h2.timed(h2.step_resident, 20, 4)
            ms_e = h2.timed(h2.step_e2e, 20, 4)
            batches[f"batch_{b}"] = dict(images_per_second=b * 20 / (ms_b / 1e3), ms_per_step=ms_b / 20,
                                         e2e_images_per_second=b * 20 / (ms_e / 1e3), e2e_ms_per_step=ms_e / 20,
                                         launches_per_step=next(g.launches for k, g in seg._graphs.items() if not k[2]))
            del h2
    del hs
    seg._graphs.clear()
    torch.cuda.empty_cache()

    city = None
    if args.city_batch > 0:       # BASELINE.json config 4, batch-sharded like the headline
        cseg = build("cityscapes", dev)
        cb_ = args.city_batch
        hc = Harness(cseg, cb_, CITY_H, CITY_W, dev, dist, 3000 + rank)
        csteps = max(4, args.steps // 2)
        ms_v = hc.timed(lambda i: hc.step_resident(i, labels=True), csteps, 3)
        ms_e = hc.timed(hc.step_e2e, csteps, 3)
        city = dict(workload="Spike2Former SDTv2 + DCN pixel decoder, Cityscapes-shape 1024x2048 (19 classes), fused argmax labels",
                    batch_per_gpu=cb_, steps=csteps, images_per_second=world * cb_ * csteps / (ms_v / 1e3), ms_per_step=ms_v / csteps,
                    e2e={"value": world * cb_ * csteps / (ms_e / 1e3), "unit": "images/s",
                         "h2d_bytes_per_step": cb_ * 3 * CITY_H * CITY_W, "d2h_bytes_per_step": cb_ * CITY_H * CITY_W})
        if rank == 0:
            croof = engine.profile_dominant(cseg, hc.devin[0], steps=2, labels=True)
            cpeak = pk["bf16_sustained"] if croof["bound"] == "tensor" else pk["hbm"]
            city["roofline"] = {"bound": croof["bound"], "kernel": croof["kernel"], "achieved": croof["achieved"],
                                "achieved_executed": croof["achieved_executed"], "peak": cpeak, "unit": croof["unit"],
                                "frac": croof["achieved"] / cpeak, "share_of_step": croof["share"],
                                "classes": class_rooflines(croof, pk, tp)}
        del hc, cseg
        torch.cuda.empty_cache()

    training = training_step_bench(dev, world, rank, dist, args.train_batch, args.train_precision) if args.train_batch > 0 else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    cfg_out = workload_config(B, world)
    out = {
        "metric": "images_per_second", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": roof["dtype"] if roof else "int8xf32", "data": "synthetic",
        "config": cfg_out,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * H * W,
                "d2h_bytes_per_step": B * H * W},
        "gpu_launches": int(launches),
        "model_tflops": value * ALG_GFLOP_PER_IMG / 1e3,
    }
    traffic = None
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    if roof and roof["kernel"] == "gemm_tc" and os.path.exists(tj):
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu launch list (mean over the
        # GEMM launches of one forward), scaled from the profiled batch to this run's batch
        t = json.load(open(tj))["gemm_i8_tc_kernel<3>"]
        traffic = t["dram_bytes_per_launch"] * B / t["batch"]
    if roof:
        peak = pk["bf16_sustained"] if roof["bound"] == "tensor" else pk["hbm"]
        out["roofline"] = {"bound": roof["bound"], "achieved": roof["achieved"], "peak": peak, "unit": roof["unit"],
                           "frac": roof["achieved"] / peak, "traffic": traffic, "kernel": roof["kernel"],
                           "flops_counted": "algorithmic: the reference layers' MACs (RepConv = 2*C*C + 9*C per pixel, unpadded channels)",
                           "achieved_executed": roof["achieved_executed"],
                           "share_of_step": roof["share"], "launches_per_step": roof["launches"],
                           "peak_source": pk["src"] + (" bf16 sustained (inside a long step)" if roof["bound"] == "tensor" else " copy"),
                           "per_class_ms": roof["per_class_ms"], "classes": class_rooflines(roof, pk, tp),
                           "whole_step": {"ms_under_events": roof["step_ms"],
                                          "model_tflops_algorithmic": value * ALG_GFLOP_PER_IMG / 1e3,
                                          "frac_of_bf16_sustained": value * ALG_GFLOP_PER_IMG / 1e3 / pk["bf16_sustained"]}}
    if tp:
        out["tensor_peaks"] = tp
    if city:
        out["cityscapes_1024x2048"] = city
    if batches:
        out["batches"] = batches
    if training:
        out["training_step"] = training
    if micro:
        out["kernels"] = micro
    if extras and not args.no_cpu_baseline:
        cb, _ = cpu_reference_run(6, 1)
        out["cpu_baseline"] = cb
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
