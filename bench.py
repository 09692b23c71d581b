#!/usr/bin/env python
"""Benchmark of the IPSR / CSA patch-shift layer (BASELINE.json metric: shift-layer fwd+bwd images/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one forward + backward of the shift layer over one batch of synthetic features.
Workload at N=1: BASELINE.json configs[1] -- batch 16, 256^2 images, 32x32x256 feature maps, centre
mask (M = 256 of N = 1024 positions masked).  With N > 1 GPUs every rank runs the same per-GPU batch on
its own synthetic shard (batch sharding: images are independent, no collective in the data path) and
`value` is the aggregate images/s -- weak scaling.

Printed JSON line (rank 0): value = device-resident throughput (CUDA events, max over ranks);
e2e = the same metric through the reference-shaped module API with HOST (pinned) buffers, H2D and D2H
copies inside the timed region; roofline = the correlation kernel against the measured bf16 peak;
cpu_baseline = the numpy port of the reference's CPU path timed on this box's host cores.
`--impl reference` times that CPU path alone (the reference is pure Python/PyTorch-on-CPU for this
layer and cannot travel to the GPU box; the pinned numpy port of it in oracle/ stands in).
"""
import argparse
import json
import math
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = dict(name="configs[1]: shift layer fwd+bwd, batch 16 per GPU, 256^2 images, 32x32x256 features, centre mask",
                B=16, C=256, H=32, W=32)
METRIC = "shift_layer_fwd_bwd_images_per_sec"
UNIT = "images/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=WORKLOAD["B"])
    ap.add_argument("--channels", type=int, default=WORKLOAD["C"])
    ap.add_argument("--size", type=int, default=WORKLOAD["H"], help="feature map side (32 = 256^2 image, 64 = 512^2)")
    ap.add_argument("--mode", default="auto", choices=["auto", "tensor", "exact"])
    ap.add_argument("--graphs", type=int, default=1, help="replay the step from a CUDA graph (1) or launch eagerly (0)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer leg (0: min(steps, 200))")
    ap.add_argument("--cpu-images", type=int, default=0, help="images of the CPU-baseline sample (0: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=1234, help="synthetic data seed (rank r uses seed + r)")
    ap.add_argument("--workload", default="shift", choices=["shift", "patch3x3", "generator"],
                    help="shift: configs[1] / configs[2], fwd+bwd of the 1x1 layer (the default line); patch3x3: configs[3], the "
                         "reference-guided forward with 3x3 patches on a 64x64x256 map (forward only: the reference has no backward "
                         "for shift_sz != 1); generator: configs[4], the full generator training step (rough + refinement U-Net in "
                         "bf16 around the shift layer, per-sample free-form masks, Adam; DDP with several GPUs)")
    ap.add_argument("--gen-batch", type=int, default=32, help="images per GPU and step of the generator workload")
    ap.add_argument("--shard", default="batch", choices=["batch", "bank"],
                    help="with several GPUs: split the batch (no data-path collective) or, for patch3x3, split the patch BANK "
                         "across the ranks and merge the (max, idx) keys with one NCCL all-reduce MAX")
    ap.add_argument("--patch-batch", type=int, default=2, help="images per step of the patch3x3 workload")
    ap.add_argument("--no-also", action="store_true",
                    help="skip the short configs[2] (512^2: batch 64, 64x64x256) run appended to the default line as 'also'")
    return ap.parse_args()


def centre_flag(H, W):
    import numpy as np
    f = np.zeros((H, W), np.int64)
    f[H // 4:3 * H // 4, W // 4:3 * W // 4] = 1
    return f.reshape(-1)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]),
                    tf_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# the reference's CPU path (numpy port pinned against the reference's own outputs, oracle/)
# ------------------------------------------------------------------------------------------------
def _host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_path_images_per_sec(C, H, W, n_images, seed=1234, kind=None):
    """fwd+bwd images/s of the reference's CPU path on this box's host cores, one image per call (the reference
    loops over the batch, models/IPSRFunction.py:46).  kind "reference": the reference's own unmodified PyTorch code
    staged under oracle/_ref (oracle/build_ref.py); kind "port": the numpy restatement oracle/ipsr_oracle.py (pinned on
    the reference's golden outputs) when the staged copy did not travel.  Returns (images/s, threads, seconds, kind)."""
    import numpy as np
    from oracle import ref_runner
    if kind is None:
        kind = "reference" if ref_runner.available() else "port"
    rng = np.random.default_rng(seed)
    flag = centre_flag(H, W)

    def inputs():
        x = rng.standard_normal((1, C, H, W)).astype(np.float32)
        ref = (np.maximum(rng.standard_normal((1, C, H, W)), 0) * 3).astype(np.float32)
        g = rng.standard_normal((1, C, H, W)).astype(np.float32)
        return x, ref, g

    if kind == "reference":
        import torch
        threads = _host_threads()
        torch.set_num_threads(threads)               # torchrun exports OMP_NUM_THREADS=1; this arm is ONE process
        S = H * 8
        mg = torch.zeros(1, 1, S, S, dtype=torch.bool)
        mg[:, :, S // 4:3 * S // 4, S // 4:3 * S // 4] = True

        def one():
            x, ref, g = (torch.from_numpy(a) for a in inputs())
            t0 = time.perf_counter()
            ref_runner.shift_fwd_bwd(x, ref, g, mg)
            return time.perf_counter() - t0
    else:
        from oracle import ipsr_oracle as O
        try:
            from threadpoolctl import threadpool_info
            threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
        except Exception:
            threads = os.cpu_count() or 1

        def one():
            x, ref, g = inputs()
            t0 = time.perf_counter()
            res = O.shift_forward(x, ref, flag, np.float32, keep_attn=True, with_gap=False)
            O.shift_backward(g, res.attn_trunc, 1.0)
            return time.perf_counter() - t0

    one()                                       # warm-up (thread pools, lazy imports, page faults)
    total = sum(one() for _ in range(n_images))
    return n_images / total, threads, total, kind


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores, same metric and
    config as our arm.  Each step is a bounded sample of the batch (the reference processes images one by one)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is one process and may use every host thread
    pool = None
    try:
        import numpy  # noqa: F401  (the BLAS pool must be loaded before it can be resized)
        from threadpoolctl import threadpool_limits
        pool = threadpool_limits(limits=_host_threads())
    except Exception as exc:                                       # keep the launcher's setting
        sys.stderr.write("threadpoolctl unavailable (%s): BLAS threads as configured by the environment\n" % exc)
    if args.workload == "generator":
        base = cpu_generator_baseline(max(1, min(args.steps, 3)))
        if base is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not staged on this box (python oracle/build_ref.py)"}), flush=True)
            return
        line = {"impl": "reference", "metric": GEN_METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": max(1, min(args.steps, 3)),
                "warmup": 1, "ms_per_step": 1e3 / base["value"] * args.gen_batch, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": gen_config(args.gen_batch, 1, False), "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return
    if args.workload == "patch3x3":
        C, H = args.channels, 64 if args.size == WORKLOAD["H"] else args.size
        steps = max(1, min(args.steps, 2))
        cpu_patch_images_per_sec(C, H, 3, 1)
        ips, threads, secs = cpu_patch_images_per_sec(C, H, 3, steps)
        line = {"impl": "reference", "metric": PATCH_METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": 1,
                "ms_per_step": 1e3 * secs / steps * args.patch_batch, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": patch_config(args.patch_batch, C, H, 3, 1, "batch"),
                "cpu_baseline": {"value": ips, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": "%d steps x 1 image, forward, fp32 (numpy port: the reference fails after computing the output for shift_sz = 3)" % steps},
                "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return
    C, H = args.channels, args.size
    per_step = max(1, min(args.batch, 2 if H <= 32 else 1))          # bounded sample of the batch per step
    steps, warm = max(1, min(args.steps, 8)), max(1, min(args.warmup, 1))
    kind = "port"
    for _ in range(warm):
        kind = cpu_path_images_per_sec(C, H, H, 1)[3]
    t_imgs, t_secs, threads = 0, 0.0, 1
    for _ in range(steps):
        ips, threads, secs, kind = cpu_path_images_per_sec(C, H, H, per_step)
        t_imgs += per_step
        t_secs += secs
    value = t_imgs / t_secs
    cfg = workload_config(args.batch, C, H, 1, args.mode)
    cfg["note"] = ("reference arm = the reference's own unmodified PyTorch implementation (oracle/_ref, staged by oracle/build_ref.py) "
                   "on the host CPU" if kind == "reference" else
                   "reference arm = numpy port (oracle/ipsr_oracle.py) pinned on the reference's golden outputs; oracle/_ref "
                   "was not staged on this box")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * t_secs / steps * (args.batch / per_step), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": "%d steps x %d image(s) of the batch, fwd+bwd, fp32" % (steps, per_step)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    del pool


def workload_config(B, C, H, world, mode):
    """The `config` object both arms print (same keys, so the driver can compare them)."""
    N = H * H
    M = int(centre_flag(H, H).sum())
    return {"workload": WORKLOAD["name"] if (B, C, H) == (16, 256, 32) else
            "shift layer fwd+bwd, batch %d per GPU, %dx%dx%d features, centre mask" % (B, H, H, C),
            "batch_per_gpu": B, "global_batch": B * world, "C": C, "H": H, "W": H, "N": N, "M": M,
            "parallelism": "batch-sharded x%d, no data-path collective" % world, "mode": mode}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def measure(args, B, C, H, K, W_, world, rank, local, dev, dist, pk, with_cpu, e2e_seconds):
    """One workload (batch B per GPU, C channels, H x H map, centre mask) on this rank; collective-free except for the
    barriers / MAX reductions of the timing contract.  Returns the fields of the JSON line (rank 0) or None."""
    import collections
    import torch
    from deepinpainting_b200 import shift_ops
    from deepinpainting_b200 import _lib as L
    from deepinpainting_b200.graphed import GraphedShiftStep
    from deepinpainting_b200.models import IPSR_model

    N = H * H
    flag = centre_flag(H, H)
    M = int(flag.sum())
    mi = shift_ops.mask_index_from_flag(torch.from_numpy(flag), dev)

    # ---- synthetic inputs: a rotating pool larger than L2 so every step starts from HBM ----
    gen = torch.Generator(device="cpu").manual_seed(args.seed + rank)
    bytes_per_set = 3 * B * C * N * 4
    pool = max(2, -(-2 * 126 * (1 << 20) // bytes_per_set) + 1)
    pool = min(pool, 16)
    sets = []
    for _ in range(pool):
        x = torch.randn(B, C, H, H, generator=gen)
        ref = torch.relu(torch.randn(B, C, H, H, generator=gen)) * 3
        g = torch.randn(B, C, H, H, generator=gen)
        sets.append((x.to(dev), ref.to(dev), g.to(dev)))

    def step_eager(i, events=None):
        x, ref, g = sets[i % pool]
        out, saved = shift_ops.shift_forward(x, ref, mi, need_grad=True, events=events)
        gin = shift_ops.shift_backward(g, saved, 1.0)
        return out, gin

    launches_per_step = shift_ops.launches_per_step(C, N, M, need_grad=True, mode=args.mode, backward=True, B=B)

    # ---- optional CUDA graphs: one per input set ----
    graphs = None
    if args.graphs:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(min(3, pool)):
                    step_eager(i)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graphs, keep = [], []
            for i in range(pool):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    keep.append(step_eager(i))
                graphs.append(gr)
        except Exception as exc:                                   # graphs are an optimisation, eager is the same work
            sys.stderr.write("CUDA graph capture failed (%s); running eagerly\n" % exc)
            graphs = None
            torch.cuda.synchronize()

    def step(i):
        if graphs is not None:
            graphs[i % pool].replay()
        else:
            step_eager(i)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world == 1:
            return [float(v)]
        outl = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(outl, t)
        return [float(o.item()) for o in outl]

    # ---- warm-up, then EXACTLY K timed steps ----
    for i in range(W_):
        step(i)
    sync_all()
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    e0.record()
    for i in range(K):
        step(W_ + i)
    e1.record()
    sync_all()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    per_rank_ms = gather_ranks(ms / K)
    ms_max = max_over_ranks(ms)
    value = world * B * K / (ms_max * 1e-3)

    # ---- live roofline of the correlation kernel: event pairs recorded by the library around it ----
    RK = min(K, 200)
    pairs = []
    for i in range(RK):
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); b_.record()                                   # materialise the handles
        pairs.append((a, b_))
    torch.cuda.synchronize()
    for i in range(RK):
        step_eager(i, events=pairs[i])
    torch.cuda.synchronize()
    corr_ms = statistics.mean(p[0].elapsed_time(p[1]) for p in pairs)
    tensor_mode = args.mode == "tensor" or (args.mode == "auto" and C % 64 == 0 and N % 128 == 0)
    cascade = tensor_mode and L.load().ipsr_tensor_cascade(B, C, N) == 1
    full_passes = L.load().ipsr_tensor_full_passes(B, C, N) if tensor_mode else 1
    flops = 2.0 * N * N * C * B                                    # algorithmic: one N x N x C correlation per image
    achieved = flops / (corr_ms * 1e-3) / 1e12
    # DRAM bytes per launch: NOT measured in this run -- read from the committed ncu --set full capture of the same
    # command line (profiles/), the source is named in the line
    traffic, traffic_src, tj_all = None, None, {}
    for fname, shape, pick in (("roofline_traffic.json", (WORKLOAD["B"], WORKLOAD["C"], WORKLOAD["H"]), "corr_tc"),
                               ("roofline_traffic_configB.json", (64, 256, 64), "corr_tc_kernel<1, 1,")):   # the single pass
        tpath = os.path.join(ROOT, "profiles", fname)
        if os.path.exists(tpath) and (B, C, H) == shape:
            with open(tpath) as fh:
                tj_all = json.load(fh)
            if tensor_mode:
                traffic = next((v for k, v in tj_all.items() if pick in k), None)
                traffic_src = "profiles/" + fname + " (ncu --set full capture of this command, committed; not re-measured in this run)"
    peak_tf = pk["tf_burst"]                                      # the kernel is timed alone by its own event pair: burst peak
    kname = "corr_fp32_kernel (FFMA)"
    if tensor_mode:
        kname = ("corr_tc_kernel pass 1 (tcgen05 fp16 hi*hi over every row; ambiguous rows are redone by the 3-pass split)" if cascade
                 else "corr_tc_kernel, one tcgen05 fp16 hi*hi pass over every row with the per-row decision in its epilogue (rows inside the rigorous error band: two exact fp32 dot products, or the exact fp32 correlation)" if full_passes == 1
                 else "corr_tc_kernel (tcgen05 fp16, 3-pass split hi*lo + lo*hi + hi*hi over every row: ceiling = 1/3 of peak)")
    roofline = {"bound": "tensor", "kernel": kname,
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "frac_of_sustained_peak": achieved / pk["tf_sustained"],
                "traffic": traffic, "traffic_source": traffic_src, "kernel_ms": corr_ms, "share_of_step": corr_ms / (ms_max / K),
                "issued_passes": full_passes,
                "tensor_pipe_utilisation": (achieved * full_passes) / peak_tf,
                "peak_source": pk["source"] + " bf16 dense, burst (kernel timed alone by its own event pair)",
                "note": "achieved = algorithmic FLOPs (2*N^2*C per image) / time of the timed correlation launch (event pair recorded by the library around it); tensor_pipe_utilisation counts the MMA passes actually issued (3 for the split over every row that small problems run, 1 for pass 1 of the cascade)"}

    # ---- workload diagnostics (one eager step): rows deferred to the exact path, attention entries that survive the
    # reference's int64 truncation ("exceptions" of the backward), hub columns
    _, dsaved = shift_ops.shift_forward(sets[0][0], sets[0][1], mi, need_grad=True, diagnostics=True)
    torch.cuda.synchronize()
    rp = dsaved.route_ptr.long()
    et = dsaved.exc_total
    diag = {"recheck_rows_per_image": float(dsaved.nrecheck.float().mean()),
            "exceptions_per_image_mean": float(et.float().mean()) if et is not None else 0.0,
            "exceptions_per_image_max": int(et.max()) if et is not None else 0,
            "exception_pool_entries": int(dsaved.exc_cap),
            "images_replaying": int((et >= shift_ops.EXC_REPLAY).sum()) if et is not None else 0,
            "three_pass_rows_per_image": float(dsaved.npass2.float().mean()),
            "exception_columns_per_image": float((dsaved.exc_cnt > 0).float().sum(1).mean()) if dsaved.exc_cnt is not None else 0.0,
            "max_routes_per_column": int((rp[:, 1:] - rp[:, :-1]).max())}

    # ---- the bandwidth-bound kernels alone, through their own C-ABI entry points (CUDA events, rotating inputs) ----
    def time_call(fn, reps=30):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        a_, b__ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for i in range(reps):
            fn(i)
        b__.record()
        torch.cuda.synchronize()
        return a_.elapsed_time(b__) / reps

    st = torch.cuda.current_stream().cuda_stream
    f32 = dict(dtype=torch.float32, device=dev)
    k_inv, k_rn = torch.empty(B, N, **f32), torch.empty(B, N, **f32)
    k_rs, k_re = torch.empty(B, N, **f32), torch.empty(B, N, **f32)
    k_xmax = torch.zeros(B, **f32)
    k_xt = torch.empty(B, N, C, **f32)
    k_rm = torch.empty(B, max(M, 1), C, **f32)
    k_nf = torch.zeros(B, dtype=torch.int32, device=dev)
    tiles_ok = (C % 64 == 0 and N % 128 == 0)
    k_xtl = torch.empty(B * C * N * 4, dtype=torch.uint8, device=dev) if tiles_ok else None
    k_rtl = torch.empty(B * C * N * 4, dtype=torch.uint8, device=dev) if tiles_ok else None

    def run_prep(i):
        x_, r_, _ = sets[i % pool]
        L.call("ipsr_extract_normalize", x_.data_ptr(), r_.data_ptr(), B, C, N, mi.rank.data_ptr(), M, k_inv.data_ptr(),
               k_rn.data_ptr(), k_xt.data_ptr(), k_rm.data_ptr(), k_xtl.data_ptr() if tiles_ok else None,
               k_rtl.data_ptr() if tiles_ok else None, k_nf.data_ptr(), k_rs.data_ptr(), k_re.data_ptr(), None,
               k_xmax.data_ptr(), st)

    Mp = -(-M // 8) * 8
    k_y = torch.randn(B, C, max(Mp, 8), **f32)
    k_out = torch.empty(B, C, H, H, **f32)

    def run_paste(i):
        x_ = sets[i % pool][0]
        L.call("ipsr_paste", x_.data_ptr(), k_y.data_ptr(), dsaved.ind.data_ptr(), mi.rank.data_ptr(), B, C, N, M,
               k_out.data_ptr(), st)

    def run_bwd(i):
        shift_ops.shift_backward(sets[i % pool][2], dsaved, 1.0)

    t_prep, t_paste, t_bwd = time_call(run_prep), time_call(run_paste), time_call(run_bwd)
    nc4 = B * N * C * 4
    kernels = []
    # algorithmic bytes per SURVEY 8(d): (a) reads x and ref (8NC) and writes their fp16 hi/lo operand images (8NC);
    # (d), (e) read one map and write one
    for name, ms_, byts, moved in (("prep_kernel (a)", t_prep, 4 * nc4, 2 * nc4 + nc4 + (2 * nc4 if tiles_ok else 0) + B * M * C * 4),
                                   ("paste_kernel (d)", t_paste, 2 * nc4, 2 * nc4),
                                   ("shift_bwd_kernel (e)", t_bwd, 2 * nc4, 2 * nc4)):
        gbs = byts / (ms_ * 1e-3) / 1e9
        kernels.append({"kernel": name, "bound": "hbm", "ms": ms_, "algorithmic_bytes": byts, "bytes_moved": moved, "achieved": gbs,
                        "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                        "traffic": next((v for k, v in tj_all.items() if name.split("_kernel")[0] in k), None)})

    # ---- e2e: the module API with HOST (pinned) buffers, H2D + D2H inside the timed region ----
    # Two pinned host buffer sets and two device buffer sets: copy-in of step i+1, compute of step i and copy-out of step
    # i-1 overlap on three streams (H2D and D2H use separate copy engines).  Three legs over the same buffers:
    #   graphed  IPSR_model forward + autograd backward captured once per buffer set (deepinpainting_b200.graphed) -- e2e.value
    #   eager    the same through plain module calls (Python + ~12 launches per step)
    #   copies   the H2D / D2H copies alone: the ceiling the link (and the host memory system) sets
    Ref = collections.namedtuple("Ref", ["relu4_3"])
    S = H * 8
    mg = torch.zeros(1, 1, S, S, dtype=torch.bool)
    mg[:, :, S // 4:3 * S // 4, S // 4:3 * S // 4] = True
    layer = IPSR_model(5 / 16.0, 1, 1, 1, 1, 1)
    layer.set_mask(mg.to(dev), 3, 5 / 16.0)
    hin = [[torch.empty(B, C, H, H).pin_memory() for _ in range(3)] for _ in range(2)]
    hout = [[torch.empty(B, C, H, H).pin_memory() for _ in range(2)] for _ in range(2)]
    for k in range(2):
        for j in range(3):
            hin[k][j].copy_(sets[k % pool][j])
    steppers = None
    if args.graphs:
        try:
            steppers = [GraphedShiftStep(layer, B, C, H, H, dev) for _ in range(2)]
        except Exception as exc:
            sys.stderr.write("graphed module path unavailable (%s)\n" % exc)
    din = [[torch.empty(B, C, H, H, device=dev) for _ in range(3)] for _ in range(2)]
    s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_cmp = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    hold = [None, None]

    def e2e_step(i, leg):
        k = i & 1
        dst = (steppers[k].x, steppers[k].ref, steppers[k].g) if leg == "graphed" else din[k]
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_cmp[k])                      # device inputs of step i-2 are no longer read
            for j in range(3):
                dst[j].copy_(hin[k][j], non_blocking=True)
            ev_in[k].record(s_in)
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(ev_in[k])
            if leg == "graphed":
                s_cmp.wait_event(ev_out[k])                 # the static outputs of step i-2 have been copied out
                yd, gd = steppers[k].replay()
            elif leg == "eager":
                xin = din[k][0].detach().requires_grad_(True)
                layer.set_ref(Ref(din[k][1]))
                y = layer(xin)
                y.backward(din[k][2])
                yd, gd = y.detach(), xin.grad
                hold[k] = (y, xin)
            else:
                yd, gd = din[k][0], din[k][2]               # copies only
            ev_cmp[k].record(s_cmp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cmp[k])
            s_out.wait_event(ev_out[k])                     # host outputs of step i-2 have been written
            if leg == "eager":
                yd.record_stream(s_out)
                gd.record_stream(s_out)
            hout[k][0].copy_(yd, non_blocking=True)
            hout[k][1].copy_(gd, non_blocking=True)
            ev_out[k].record(s_out)

    def e2e_leg(leg, steps):
        for i in range(4):
            e2e_step(i, leg)
        sync_all()
        t0 = time.perf_counter()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(s_in)                                     # first operation of the region: the first H2D copy
        for i in range(steps):
            e2e_step(i, leg)
        f1.record(s_out)                                    # last operation: the last D2H copy has landed
        enqueue = time.perf_counter() - t0                  # host time to ENQUEUE the region (Python + launches)
        sync_all()
        wall = time.perf_counter() - t0
        secs = max_over_ranks(f0.elapsed_time(f1) * 1e-3)
        return {"value": world * B * steps / secs, "steps": steps, "region_s": secs, "wall_s": wall,
                "host_enqueue_ms_per_step": max_over_ranks(enqueue / steps * 1e3)}

    main_leg = "graphed" if steppers is not None else "eager"
    pilot = e2e_leg(main_leg, 8)
    EK = args.e2e_steps or int(min(5000, max(20, math.ceil(e2e_seconds * 8 / max(pilot["region_s"], 1e-6)))))
    legs = {main_leg: e2e_leg(main_leg, EK)}
    other = "eager" if main_leg == "graphed" else None
    if other:
        legs[other] = e2e_leg(other, max(20, EK // 4))
    legs["copies_only"] = e2e_leg("copies", max(20, EK // 4))
    h2d, d2h = 3 * B * C * N * 4, 2 * B * C * N * 4
    e2e = {"value": legs[main_leg]["value"], "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": legs[main_leg]["steps"], "region_s": legs[main_leg]["region_s"], "wall_s": legs[main_leg]["wall_s"],
           "host_enqueue_ms_per_step": legs[main_leg]["host_enqueue_ms_per_step"],
           "api": ("deepinpainting_b200.graphed.GraphedShiftStep: IPSR_model.forward + autograd backward captured into a CUDA graph per buffer set"
                   if main_leg == "graphed" else "IPSR_model.forward + autograd backward (eager)") + ", pinned host buffers",
           "eager_module_api": legs.get("eager"),
           "copies_only_ceiling": dict(legs["copies_only"], gbs_h2d=legs["copies_only"]["value"] / (world * B) * h2d / 1e9,
                                       gbs_d2h=legs["copies_only"]["value"] / (world * B) * d2h / 1e9,
                                       note="the same H2D / D2H copies on the same three streams with the layer left out: the ceiling set by the PCIe link and, with several ranks, by the host memory system they share"),
           "timing": "CUDA events: first H2D copy -> last D2H copy of the region, max over ranks (3 streams overlap H2D / compute / D2H)"}

    # ---- CPU baseline (rank 0, N=1 only) ----
    cpu = None
    if with_cpu and rank == 0 and world == 1:
        from oracle import ref_runner
        real = ref_runner.available()
        n_img = args.cpu_images or ((40 if real else 1200) if H <= 32 else (6 if real else 40))   # ~10-20 s of CPU work
        ips, threads, secs, kind = cpu_path_images_per_sec(C, H, H, n_img)
        cpu = {"value": ips, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": "%d images of the workload (B=1 each: the reference loops over the batch), fwd+bwd, fp32, %.1f s" % (n_img, secs)}

    cfg = workload_config(B, C, H, world, args.mode)
    cfg.update({"cuda_graphs": graphs is not None,
                "correlation": ("fp32 results; tcgen05 fp16 hi/lo split with fp32 accumulation ("
                                + ("precision cascade: 1 pass, 3-pass split on the ambiguous rows" if cascade
                                   else "1 pass, rows inside the rigorous error band settled in exact fp32" if full_passes == 1
                                   else "3-pass split over every row") + "), ties and ambiguous rows resolved in exact fp32")
                               if tensor_mode else "fp32 FFMA",
                "l2": "rotating pool of %d input sets (%d MB) > 126 MB L2" % (pool, pool * bytes_per_set >> 20)})
    del sets, graphs, steppers, hin, hout, din
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
            "ms_per_step": ms_max / K, "per_rank_ms_per_step": per_rank_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "diagnostics": diag, "kernels": kernels,
            "gpu_launches": launches_per_step * K}


PATCH_METRIC = "patch3x3_forward_images_per_sec"
GEN_METRIC = "generator_training_step_images_per_sec"


def freeform_masks(B, S, seed):
    """One free-form mask per sample: a few random rectangles (bool [B,1,S,S], True = hole)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    m = np.zeros((B, 1, S, S), bool)
    for b in range(B):
        for _ in range(4):
            y, x = rng.integers(0, S - S // 4, 2)
            h, w = rng.integers(S // 16, S // 3, 2)
            m[b, 0, y:y + h, x:x + w] = True
    return m


def gen_config(B, world, bf16):
    return {"workload": "configs[4]: full generator training step -- rough U-Net netP + refinement U-Net netG (shift layer, InnerCos, "
                        "InnerCos2 at 32x32x512) + VGG-16 relu4_3 of reference and ground truth, 256^2 images, one free-form mask per "
                        "sample, L1 losses, Adam; discriminators / GAN loss out of scope",
            "batch_per_gpu": B, "global_batch": B * world, "image": 256, "convs": "cuDNN " + ("bf16 autocast, channels_last" if bf16 else "fp32"),
            "shift_layer": "fp32 results (tcgen05 fp16 split + exact recheck), per-sample masks in one batched call",
            "parallelism": "DDP x%d (NCCL gradient all-reduce overlapped with backward); the shift layer needs no collective" % world}


def measure_generator(args, world, rank, local, dev, dist, pk):
    import torch
    from deepinpainting_b200 import generator as G
    B, S = args.gen_batch, 256
    step = G.GeneratorStep(dev, bf16=True, ddp=world > 1, seed=args.seed)
    gen = torch.Generator(device="cpu").manual_seed(args.seed + rank)
    pool = 3
    batches = []
    for i in range(pool):
        img = torch.rand(B, 3, S, S, generator=gen) * 2 - 1
        ref = torch.rand(B, 3, S, S, generator=gen) * 2 - 1
        mask = torch.from_numpy(freeform_masks(B, S, args.seed + 100 * rank + i))
        batches.append((img, mask, ref))
    dbatches = [(a.to(dev), m.to(dev), r.to(dev)) for a, m, r in batches]
    # time spent inside the shift layer (forward) per step, by events around the module
    layer = step.shift_layers[0]
    marks = []
    layer.register_forward_pre_hook(lambda m, i: marks.append([torch.cuda.Event(enable_timing=True), None]) or marks[-1][0].record())

    def post(m, i, o):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks[-1][1] = ev
    layer.register_forward_hook(post)

    def one(i, host=False):
        if host:
            a, m, r = (t.pin_memory() if not t.is_pinned() else t for t in batches[i % pool])
            a, m, r = a.to(dev, non_blocking=True), m.to(dev, non_blocking=True), r.to(dev, non_blocking=True)
        else:
            a, m, r = dbatches[i % pool]
        step.set_input(a, m, r)
        return step.optimize_parameters()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    W_, K = 6, max(10, min(args.steps, 30))               # the first steps pay cuDNN's algorithm selection and the allocator's growth
    for i in range(W_):
        one(i)
    sync_all()
    marks.clear()
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    e0.record()
    for i in range(K):
        one(W_ + i)
    e1.record()
    sync_all()
    clocks = sampler.stop()
    ms_max = max_over_ranks(e0.elapsed_time(e1))
    shift_fwd_ms = statistics.mean(a.elapsed_time(b) for a, b in marks if b is not None) if marks else None
    value = world * B * K / (ms_max * 1e-3)
    # e2e: pinned host batches in, the loss value out, every step
    pinned = [tuple(t.pin_memory() for t in b_) for b_ in batches]
    batches[:] = pinned
    for i in range(2):
        float(one(i, host=True))
    sync_all()
    EK = max(5, K // 2)
    t0 = time.perf_counter()
    for i in range(EK):
        float(one(i, host=True))                          # .item(): the D2H read of the step's result
    sync_all()
    wall = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": world * B * EK / wall, "unit": UNIT, "h2d_bytes_per_step": B * (2 * 3 * S * S * 4 + S * S), "d2h_bytes_per_step": 4,
           "steps": EK, "api": "deepinpainting_b200.generator.GeneratorStep.set_input + optimize_parameters, pinned host batches, loss.item() per step",
           "timing": "host wall clock around the region (every step ends in a device -> host read), max over ranks"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_generator_baseline()
    if rank != 0:
        return None
    return {"metric": GEN_METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_, "ms_per_step": ms_max / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 convolutions, f32 shift layer", "data": "synthetic",
            "config": gen_config(B, world, True), "shift_layer_forward_ms": shift_fwd_ms,
            "shift_layer_forward_share": (shift_fwd_ms / (ms_max / K)) if shift_fwd_ms else None,
            "roofline": None, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": 12 * K}


def cpu_generator_baseline(steps=2):
    """The reference's own generator (models/networks.py, unmodified, staged under oracle/_ref) with the reference's own
    shift layer on the host CPU: forward + backward of netP and netG on ONE image per step (the reference's batch size)."""
    import torch
    from oracle import ref_runner
    if not ref_runner.available():
        return None
    threads = _host_threads()
    torch.set_num_threads(threads)
    ref_runner.modules()
    root = ref_runner.reference_dir()
    with ref_runner.cpu_shim():
        sys.path.insert(0, root)
        try:
            import models.networks as ref_networks
        finally:
            sys.path.remove(root)

        class Opt:
            threshold, fixed_mask, shift_sz, stride, mask_thred, triple_weight, strength, skip = 5 / 16.0, 1, 1, 1, 1, 1, 1, 0
        S = 256
        mask = torch.zeros(1, 1, S, S, dtype=torch.bool)
        mask[:, :, S // 4:3 * S // 4, S // 4:3 * S // 4] = True
        netG, cos, cos2, shift = ref_networks.define_G(6, 3, 64, "unet_ipsr", Opt, mask, "instance", False, "normal", [], 0.02)
        netP, _, _, _ = ref_networks.define_G(3, 3, 64, "unet_256", Opt, mask, "instance", False, "normal", [], 0.02)
        import collections
        R = collections.namedtuple("R", ["relu4_3"])
        total = 0.0
        for i in range(steps + 1):
            img = torch.rand(1, 3, S, S) * 2 - 1
            for m in shift:
                m.set_mask(mask, 3, Opt.threshold)
                m.set_ref(R(torch.relu(torch.randn(1, 512, 32, 32))))
            for m in cos + cos2:
                m.set_mask(mask, Opt)
                m.set_target(torch.relu(torch.randn(1, 512, 32, 32)))
            t0 = time.perf_counter()
            p = netP(img)
            out = netG(torch.cat([p, img], 1))
            ((out - img).abs().mean() + (p - img).abs().mean()).backward()
            if i > 0:
                total += time.perf_counter() - t0
    return {"value": steps / total, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": "%d steps x 1 image: forward + backward of the reference's netP and netG (VGG and optimiser left out), fp32, %.1f s" % (steps, total)}


def patch_config(B, C, H, k, world, shard):
    nH = H - k + 1
    return {"workload": "configs[3]: reference-guided shift layer forward, %dx%dx%d features, %dx%d patches (P = %d patch positions, "
                        "rows of K = %d), centre hole, batch %d" % (H, H, C, k, k, nH * nH, C * k * k, B),
            "batch": B, "C": C, "H": H, "W": H, "patch": k, "stride": 1, "P": nH * nH, "K": C * k * k,
            "parallelism": ("bank-sharded x%d: every rank correlates P/%d bank columns, one NCCL all-reduce MAX of the (max, idx) keys"
                            % (world, world)) if shard == "bank" and world > 1 else "batch-sharded x%d, no data-path collective" % world}


def measure_patch(args, world, rank, local, dev, dist, pk):
    """BASELINE.json configs[3]: forward of the layer with shift_sz = 3 on a 64 x 64 x 256 map (models/IPSRFunction.py:46-133
    with 3 x 3 patches; the reference has no backward for it).  --shard bank: all ranks hold the same images, the patch bank
    is split across them and the (max, idx) keys are merged by one all-reduce MAX."""
    import torch
    from deepinpainting_b200 import shift_ops
    from deepinpainting_b200.sharding import shard_bank, allreduce_max_keys
    B, C, H, k = args.patch_batch, args.channels, 64 if args.size == WORKLOAD["H"] else args.size, 3
    nH = H - k + 1
    P, K = nH * nH, C * k * k
    bank = args.shard == "bank" and world > 1
    feat = torch.zeros(H, H, dtype=torch.uint8, device=dev)
    feat[H // 4:3 * H // 4, H // 4:3 * H // 4] = 1
    mi = shift_ops.build_flags(feat, k, 1, 1)
    gen = torch.Generator(device="cpu").manual_seed(args.seed + (0 if bank else rank))     # bank mode: the SAME images everywhere
    pool = 4
    sets = [(torch.randn(B, C, H, H, generator=gen).abs().to(dev), (torch.relu(torch.randn(B, C, H, H, generator=gen)) * 3).to(dev))
            for _ in range(pool)]
    Ppad = -(-P // 128) * 128
    cb, ce = (0, -1)
    if bank:
        cb, ce = shard_bank(Ppad, world, rank)
        ce = min(ce, P)
        cb = min(cb, P)
    ar_ms = []

    def reduce_keys(keys):
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        allreduce_max_keys(keys)
        b_.record()
        ar_ms.append((a, b_))

    def step(i):
        x, ref = sets[i % pool]
        return shift_ops.shift_forward_patches(x, ref, mi, k, 1, mode=args.mode if args.mode != "auto" else None,
                                               col_begin=cb, col_end=ce, reduce_max=reduce_keys if bank else None)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    W_, Ks = max(3, args.warmup), min(args.steps, 100)
    for i in range(W_):
        step(i)
    sync_all()
    ar_ms.clear()
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    e0.record()
    for i in range(Ks):
        step(W_ + i)
    e1.record()
    sync_all()
    clocks = sampler.stop()
    ms_max = max_over_ranks(e0.elapsed_time(e1))
    images = B * Ks * (1 if bank else world)
    value = images / (ms_max * 1e-3)
    allreduce_ms = statistics.mean(p[0].elapsed_time(p[1]) for p in ar_ms) if ar_ms else None
    # the correlation alone (tcgen05 three-pass split over this rank's bank columns), by its C-ABI entry points
    corr = shift_ops.time_wide_patch_correlation(sets[0][0], sets[0][1], k, 1, cb, ce if ce >= 0 else P, reps=10)
    cols = (ce if ce >= 0 else P) - cb
    flops = 2.0 * P * max(cols, 0) * K * B
    achieved = flops / (corr["gemm_ms"] * 1e-3) / 1e12 if corr["gemm_ms"] > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "corr_tc_kernel<1,3,false> (tcgen05 fp16 three-pass split, both operands streamed: K = %d)" % K,
                "achieved": achieved, "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": achieved / pk["tf_burst"],
                "issued_passes": 3, "tensor_pipe_utilisation": 3 * achieved / pk["tf_burst"], "kernel_ms": corr["gemm_ms"],
                "share_of_step": corr["gemm_ms"] / (ms_max / Ks), "traffic": None,
                "peak_source": pk["source"] + " bf16 dense, burst (kernel timed alone)",
                "note": "achieved = algorithmic FLOPs (2 * P * bank columns of this rank * K per image) / time of the GEMM launch"}
    # e2e: host pinned buffers in, result out
    hx = [t.cpu().pin_memory() for t in sets[0]]
    hout = torch.empty(B, C, H, H).pin_memory()
    dx = [torch.empty_like(t) for t in sets[0]]

    def e2e_once():
        dx[0].copy_(hx[0], non_blocking=True)
        dx[1].copy_(hx[1], non_blocking=True)
        out, _ = shift_ops.shift_forward_patches(dx[0], dx[1], mi, k, 1, col_begin=cb, col_end=ce, reduce_max=reduce_keys if bank else None)
        hout.copy_(out, non_blocking=True)

    for _ in range(2):
        e2e_once()
    sync_all()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    EK = max(5, min(Ks, 40))
    f0.record()
    for _ in range(EK):
        e2e_once()
    f1.record()
    sync_all()
    e2e_s = max_over_ranks(f0.elapsed_time(f1) * 1e-3)
    e2e = {"value": B * EK * (1 if bank else world) / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 2 * B * C * H * H * 4,
           "d2h_bytes_per_step": B * C * H * H * 4, "steps": EK, "api": "shift_ops.shift_forward_patches (what IPSRFunction.apply calls for shift_sz = 3), pinned host buffers"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ips, threads, secs = cpu_patch_images_per_sec(C, H, k, 1)
        cpu = {"value": ips, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "1 image, forward, fp32, %.1f s (numpy port: the reference itself fails after computing the output for shift_sz = 3, IPSRFunction.py:134)" % secs}
    if rank != 0:
        return None
    return {"metric": PATCH_METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": Ks, "warmup": W_, "ms_per_step": ms_max / Ks,
            "higher_is_better": True, "scaling": "strong" if bank else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": patch_config(B, C, H, k, world, args.shard), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "allreduce_ms": allreduce_ms, "allreduce_bytes": 8 * B * P if bank else 0,
            "stages_ms": corr, "gpu_launches": corr["launches_per_step"] * Ks}


def cpu_patch_images_per_sec(C, H, k, n_images, seed=1234):
    import numpy as np
    from oracle import ipsr_oracle as O
    try:
        from threadpoolctl import threadpool_info
        threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        threads = os.cpu_count() or 1
    rng = np.random.default_rng(seed)
    fm = np.zeros((H, H), np.uint8)
    fm[H // 4:3 * H // 4, H // 4:3 * H // 4] = 1
    total = 0.0
    for _ in range(n_images):
        x = np.abs(rng.standard_normal((1, C, H, H))).astype(np.float32)
        ref = (np.maximum(rng.standard_normal((1, C, H, H)), 0) * 3).astype(np.float32)
        t0 = time.perf_counter()
        O.shift_forward_patches(x, ref, fm, k, 1, 1)
        total += time.perf_counter() - t0
    return n_images / total, threads, total


def run_ours(args):
    import torch
    import torch.distributed as dist
    from deepinpainting_b200 import shift_ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    shift_ops.config["correlation_mode"] = args.mode
    pk = peaks()
    B, C, H = args.batch, args.channels, args.size
    if args.workload in ("patch3x3", "generator"):
        line = (measure_patch if args.workload == "patch3x3" else measure_generator)(args, world, rank, local, dev, dist, pk)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    line = measure(args, B, C, H, args.steps, max(3, args.warmup), world, rank, local, dev, dist, pk,
                   with_cpu=not args.no_cpu_baseline, e2e_seconds=1.2)
    # ---- the metric's second size (BASELINE.json: "at 256^2/512^2 ... 1/2/4/8 B200"): configs[2] (512^2: batch 64 per GPU,
    # 64x64x256) runs on EVERY rank too, so that the scaling record carries both sizes; reported under "also"
    if not args.no_also and (B, C, H) == (WORKLOAD["B"], WORKLOAD["C"], WORKLOAD["H"]) and args.mode == "auto":
        try:
            sub = measure(args, 64, 256, 64, min(args.steps, 60), 5, world, rank, local, dev, dist, pk, with_cpu=False,
                          e2e_seconds=1.0)
            if rank == 0:
                line["also"] = [sub]
        except Exception as exc:                                   # the headline line must survive a failure of the extra leg
            sys.stderr.write("configs[2] leg failed on rank %d: %r\n" % (rank, exc))
            if world > 1:
                raise
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
