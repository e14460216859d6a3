#!/usr/bin/env python
"""bench.py -- Spike2Former 512x512 batch-sharded inference on N B200s (BASELINE.json metric: imgs/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

One "step" = one forward pass of the whole hot path (SDTv2 backbone + DCN pixel decoder + MaskFormer
head -> seg logits [B,150,512,512]) over one batch of B synthetic 512x512 images per GPU.
  value : whole-job images/s with the input batches already resident in HBM (CUDA events, max over ranks)
  e2e   : the same through the public API with HOST buffers: pinned uint8 images (the dataset format the reference's
          SegDataPreProcessor receives) -> H2D -> normalise + forward -> argmax label map (uint8) -> D2H, every step,
          inside the timed region
  roofline : the dominant kernel class (spike GEMM / conv) timed live with CUDA events on the launch stream
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
CITY_H, CITY_W = 1024, 2048    # BASELINE.json config 4: Cityscapes shape (19 classes), reported as a secondary number


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
    """The reference's CPU implementation of the path: oracle/port.py (bit-exact restatement, see tests/golden)."""
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


def kernel_microbench(dev):
    """BASELINE.json config 2 on one GPU: fused NI-LIF (B=64, N=1024, C=512, T=1) and the MLP spike GEMM 512->2048."""
    from spike2former_b200 import ops

    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, iters=10):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush.view(torch.int64).sum()                  # evict the working set, leave clean lines in L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(400000)                      # the launch is queued before the GPU reaches e0
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts) * 1e-3

    g = torch.Generator().manual_seed(0)
    x = (torch.rand(64, 1024, 512, generator=g) * 12 - 2).to(dev)
    lv = torch.empty(x.shape, dtype=torch.int8, device=dev)
    t = timed(lambda: ops.nilif(x, out=lv))
    nilif = dict(workload="NI-LIF B=64 N=1024 C=512 T=1 D=8", us=t * 1e6, gbs=x.numel() * 5 / t / 1e9,
                 algorithmic_bytes=x.numel() * 5)
    bn_s, bn_b = (torch.rand(512, generator=g) + 0.5).to(dev), torch.randn(512, generator=g).to(dev)
    t = timed(lambda: ops.nilif(x, scale=bn_s, shift=bn_b, out=lv))
    nilif["with_folded_bn_affine"] = dict(us=t * 1e6, gbs=x.numel() * 5 / t / 1e9)
    n, Hh, Ww, cin, cout = 64, 32, 32, 512, 2048
    a = torch.randint(0, 9, (n, Hh, Ww, cin), generator=g, dtype=torch.int8).to(dev)
    w = torch.randn(cout, cin, generator=g) / cin ** 0.5
    packed, rowscale = ops.pack_weights_i8(w, 1, cin, 3)
    packed, sc, sh = packed.to(dev), (rowscale / 8).to(dev), torch.zeros(cout, device=dev)
    t = timed(lambda: ops.gemm_tc(a, packed, n=n, H=Hh, W=Ww, Cin=cin, Cout=cout, scale=sc, shift=sh, pieces=3,
                                  want_spike=True))
    gemm = dict(workload="spike GEMM fc1 512->2048, 65536 tokens, int8 spikes x 3 int8 weight planes, NI-LIF epilogue",
                us=t * 1e6, tflops_algorithmic=2.0 * n * Hh * Ww * cin * cout / t / 1e12)
    return dict(nilif_cfg2=nilif, gemm_cfg2=gemm)


def cityscapes_throughput(dev, world, dist, B, steps=4):
    """Secondary number (BASELINE.json config 4): Cityscapes config at 1024x2048, batch-sharded, inputs resident."""
    import spike2former_b200 as s2f
    from spike2former_b200 import synth

    cfg = s2f.configs.cityscapes()
    seg = s2f.build_segmentor(cfg)
    seg.load_state_dict(synth.synthetic_checkpoint("cityscapes", cfg), strict=True)
    seg = seg.to(dev)
    g = torch.Generator().manual_seed(7)
    xs = [torch.randn(B, 3, CITY_H, CITY_W, generator=g).to(dev) for _ in range(2)]
    with torch.no_grad():
        for i in range(3):
            seg.predict_labels(xs[i & 1])
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            seg.predict_labels(xs[i & 1])
        e1.record()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    del seg, xs
    torch.cuda.empty_cache()
    return dict(workload="Spike2Former SDTv2 + DCN pixel decoder, Cityscapes-shape 1024x2048 (19 classes), fused argmax",
                batch_per_gpu=B, images_per_second=world * B / (ms / 1e3), ms_per_step=ms)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--city-batch", type=int, default=4, help="images per GPU for the 1024x2048 secondary number (0 = skip)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warm = max(1, min(args.steps, 12)), max(1, min(args.warmup, 2))
        cb, ms = cpu_reference_run(steps, warm)
        print(json.dumps({
            "impl": "reference", "metric": "images_per_second", "value": cb["value"], "unit": "images/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_step": 1, "device": "host CPU"},
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

    import spike2former_b200 as s2f
    from spike2former_b200 import engine, ops, synth

    cfg = s2f.configs.ade20k()
    seg = s2f.build_segmentor(cfg)
    seg.load_state_dict(synth.synthetic_checkpoint("ade20k", cfg), strict=True)
    seg = seg.to(dev)
    B = args.batch
    NBUF = 4                                   # rotate input batches: 4 x B x 3 MB > L2 for B >= 12; the per-step
    g = torch.Generator().manual_seed(1000 + rank)   # working set (GBs of activations) thrashes L2 anyway
    devin = [torch.randn(B, 3, H, W, generator=g).to(dev) for _ in range(NBUF)]      # resident fp32 batches (`value`)
    host = [torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8).pin_memory() for _ in range(NBUF)]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        with torch.no_grad():
            return seg.encode_decode(devin[i % NBUF])

    # end-to-end: pinned uint8 host batch -> H2D (copy stream, double buffered so the copy of step i+1 overlaps the
    # forward of step i) -> SegDataPreProcessor kernel + forward with fused argmax (one CUDA graph) -> D2H of the uint8
    # label map; all of it inside the timed region
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [torch.empty(B, 3, H, W, dtype=torch.uint8, device=dev) for _ in range(2)]
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    host_outs = [torch.empty(B, H, W, dtype=torch.uint8).pin_memory() for _ in range(2)]

    def step_e2e(i):
        j = i & 1
        main = torch.cuda.current_stream()
        with torch.no_grad():
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_free[j])
                stage[j].copy_(host[i % NBUF], non_blocking=True)
                ev_ready[j].record(copy_stream)
            main.wait_event(ev_ready[j])
            labels = seg.predict_labels(stage[j])
            ev_free[j].record(main)
            host_outs[j].copy_(labels, non_blocking=True)
        return host_outs[j]

    # ------------------------------------------------------------------ device-resident timing
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_resident(i)
    e1.record()
    barrier()
    launches = ops.launch_count() - l0
    if launches == 0:      # CUDA-graph replay: the library's counter only sees the capture; one replay = that many kernels
        launches = args.steps * sum(g.launches for k, g in seg._graphs.items() if not k[2])
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    t = torch.tensor([ms], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max / 1e3)

    # ------------------------------------------------------------------ end-to-end (host buffers) timing
    for i in range(4):
        step_e2e(i)
    barrier()
    e0.record()
    for i in range(args.steps):
        step_e2e(i)
    e1.record()
    barrier()                                  # every H2D, forward and D2H of the timed steps has completed here
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(t.item()) / 1e3)

    # ------------------------------------------------------------------ roofline of the dominant kernel class
    roof = engine.profile_dominant(seg, devin[0], steps=min(args.steps, 3)) if rank == 0 else None

    micro = kernel_microbench(dev) if rank == 0 else None
    city = cityscapes_throughput(dev, world, dist, args.city_batch) if args.city_batch > 0 else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    pk = peaks()
    out = {
        "metric": "images_per_second", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": roof["dtype"] if roof else "int8xf32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": world * B, "parallelism": f"dp{world}",
                   "weights": "seeded random init + shipped calibration statistics (spike2former_b200/synth.py)",
                   "l2": f"{NBUF} rotating input batches; per-step working set >> 126 MB L2",
                   "value_input": "fp32 [B,3,512,512] resident in HBM (encode_decode -> fp32 logits [B,150,512,512])",
                   "e2e_input": "uint8 [B,3,512,512] pinned host memory -> SegDataPreProcessor + predict -> uint8 labels -> host",
                   "alg_gflop_per_image": ALG_GFLOP_PER_IMG},
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
        # 179 GEMM launches of one forward), scaled from the profiled batch to this run's batch
        t = json.load(open(tj))["gemm_i8_tc_kernel<3>"]
        traffic = t["dram_bytes_per_launch"] * B / t["batch"]
    if roof:
        peak = pk["bf16_sustained"] if roof["bound"] == "tensor" else pk["hbm"]
        out["roofline"] = {"bound": roof["bound"], "achieved": roof["achieved"], "peak": peak, "unit": roof["unit"],
                           "frac": roof["achieved"] / peak, "traffic": traffic, "kernel": roof["kernel"],
                           "share_of_step": roof["share"], "launches_per_step": roof["launches"],
                           "peak_source": pk["src"] + (" bf16 sustained (inside a long step)" if roof["bound"] == "tensor" else " copy"),
                           "per_class_ms": roof["per_class_ms"]}
    if city:
        out["cityscapes_1024x2048"] = city
    if micro:
        micro["nilif_cfg2"]["frac_of_measured_hbm"] = micro["nilif_cfg2"]["gbs"] / pk["hbm"]
        micro["nilif_cfg2"]["frac_of_8tbs_nominal"] = micro["nilif_cfg2"]["gbs"] / 8000.0
        micro["nilif_cfg2"]["with_folded_bn_affine"]["frac_of_measured_hbm"] = micro["nilif_cfg2"]["with_folded_bn_affine"]["gbs"] / pk["hbm"]
        # the kernel executes 3 int8 MACs (digit planes) per algorithmic MAC; int8 dense peak: nominal 4.5 POP/s (2x the
        # nominal 2.25 PFLOP/s bf16), no int8 figure is in MEASURED_PEAKS.json.  The measured tensor-pipe activity of the
        # same launch is in the committed ncu capture.
        micro["gemm_cfg2"]["int8_ops_executed_per_s"] = 3.0 * micro["gemm_cfg2"]["tflops_algorithmic"] * 1e12
        micro["gemm_cfg2"]["frac_of_nominal_int8_4500T"] = 3.0 * micro["gemm_cfg2"]["tflops_algorithmic"] / 4500.0
        micro["gemm_cfg2"]["algorithmic_vs_measured_bf16_burst"] = micro["gemm_cfg2"]["tflops_algorithmic"] / pk["bf16"]
        micro["gemm_cfg2"]["ncu_tensor_pipe_active"] = "77.4 % (profiles/r1z_gemm_cfg2fc1_full.md)"
        out["kernels"] = micro
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_reference_run(6, 1)
        out["cpu_baseline"] = cb
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
