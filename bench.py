#!/usr/bin/env python
"""bench.py -- STF encode+decode throughput at 768x512 (BASELINE.json metric, config 3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A "step" = compress() + decompress() of one batch of B synthetic 768x512 RGB images per GPU (weak
scaling: every rank codes its own B images, no data-path collective).
One JSON line is printed by rank 0:
  value        Mpixel/s, whole job, inputs resident in HBM at the start of the timed region
  e2e          same through the public API with HOST buffers: pinned-host images -> H2D -> compress ->
               byte strings -> decompress -> x_hat -> D2H, all inside the timed region
  roofline     dominant stf_b200 kernel family (by device time over one instrumented step), classed HBM- or tensor-bound by
               its arithmetic intensity: algorithmic bytes (flops) / CUDA-event time vs the measured peak;
               roofline_by_kernel lists every family; tensor_pipe is the metric's second half
  extras       sub-records for BASELINE configs 5 (training step, NCCL all-reduce at N > 1) and 4 (WACNN 2048x1408), and the
               unmodified reference moved to the GPU (eager, batch 1) as a context number
  cpu_baseline the UNMODIFIED reference (the byte-for-byte copy `make -C oracle` leaves in the git-ignored oracle/_ref:
               its Python model code + its own C++ rANS extension) timed on this box's host cores (N=1 only)
`--impl reference` times that CPU implementation alone (rank 0 only), same metric/unit/config.
Nothing here reads /root/reference at run time on the GPU box.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 512, 768          # Kodak-size image (rows x cols)
METRIC = "STF encode+decode Mpixel/s at 768x512"


def load_peaks():
    """Roofline denominators: the driver-measured copy bandwidth and cuBLAS bf16 throughput (MEASURED_PEAKS.json), else the
    fallback of B200_PROFILING.md.  TF32 tensor peak = bf16 / 2 (the nominal ratio: 1.1 vs 2.25 PFLOP/s dense)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "tf32_tflops": float(d["bf16_tflops"]) / 2,
                "source": "measured (MEASURED_PEAKS.json: hbm_gbs; bf16_tflops burst / 2 for TF32)"}
    return {"hbm_gbs": 6650.0, "tf32_tflops": 1590.0 / 2, "source": "fallback (B200_PROFILING.md)"}


def _ncu_record():
    p = os.path.join(ROOT, "profiles", "r2_ncu_step_summary.json")
    return json.load(open(p)) if os.path.exists(p) else None


def measured_traffic(kernel, bound, launches, batch):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of a kernel family from the committed ncu capture
    of one eager step (profiles/r2_ncu_step_summary.json, written by tools/ncu_summary.py), scaled linearly to this batch."""
    d = _ncu_record()
    key = f"{kernel}|{bound}"
    if not d or key not in d.get("families", {}):
        return None, None
    k = d["families"][key]
    per_launch = (k["dram_read_bytes"] + k["dram_write_bytes"]) / k["launches"] * batch / d["batch"]
    return per_launch, f"ncu capture at batch {d['batch']} scaled x{batch / d['batch']:g} (profiles/r2_ncu_step_summary.json)"


def ncu_tensor_pipe():
    """sm__pipe_tensor_cycles_active (pct of peak sustained, time-weighted) per family from the same committed capture."""
    d = _ncu_record()
    if not d:
        return None
    return {k: v.get("tensor_pipe_pct") for k, v in d.get("families", {}).items() if v.get("tensor_pipe_pct") is not None}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_weights():
    import torch
    from stf_b200.synth import synthetic_state_dict
    spec = {k: (tuple(s), getattr(torch, d.split(".")[-1])) for k, (s, d) in
            json.load(open(os.path.join(ROOT, "tests", "golden", "stf_spec.json"))).items()}
    return synthetic_state_dict(spec, 0)


# ----------------------------------------------------------------------------- CPU reference arm

def _reference_net(device="cpu"):
    """The UNMODIFIED reference model (compressai.zoo.models['stf'] from oracle/_ref/compressai -- the byte-for-byte copy
    `make -C oracle` leaves there -- or /root/reference in the build container) with the bench's synthetic weights."""
    import torch
    from oracle.ref_import import import_reference
    import_reference()
    from compressai.zoo import models as zoo
    net = zoo["stf"]()
    torch.nn.Module.load_state_dict(net, synthetic_weights(), strict=False)
    net = net.to(device).eval()
    net.update(force=True)
    return net


def cpu_reference_rate(steps, warmup, log=None):
    """The reference's own CPU implementation of the path, timed on this box's host cores with all of them
    (torch intra-op threads = os.cpu_count()): compress() + decompress() of one 768x512 image per step.
    kind "reference": the unmodified reference Python + its own C++ rANS extension (oracle/_ref);
    kind "port": the oracle restatement (only when the reference copy was not built)."""
    import torch
    from stf_b200.synth import synthetic_image
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    kind, what = "reference", "unmodified reference (compressai.models.stf.SymmetricalTransFormer.compress/decompress, its own C++ rANS)"
    try:
        net = _reference_net("cpu")

        def run(x):
            with torch.no_grad():
                enc = net.compress(x)
                net.decompress(enc["strings"], enc["shape"])
    except Exception as e:   # the copy was not built: the oracle port stands in
        from oracle import codec as OC
        kind, what = "port", f"oracle port of stf.py on torch CPU ops (reference copy unavailable: {type(e).__name__})"
        ora = OC.StfOracle(synthetic_weights())

        def run(x):
            enc = ora.compress(x)
            ora.decompress(enc["strings"], enc["shape"])
    times = []
    for i in range(warmup + steps):
        x = synthetic_image(1, H, W, seed=100 + i)
        t0 = time.perf_counter()
        run(x)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if log:
            log(f"cpu reference step {i}: {dt:.2f}s")
    total = sum(times)
    mpx = steps * H * W / total / 1e6
    return mpx, total / steps * 1e3, {"value": mpx, "unit": "Mpixel/s", "cores": torch.get_num_threads(), "kind": kind,
                                      "sample": f"{steps} x (1 image 768x512 compress+decompress) after {warmup} warm-up, {what}"}


def gpu_eager_reference_rate(steps=3, warmup=1):
    """Context number (eval_model/__main__.py:103-110): the unmodified reference moved .to("cuda"), eager,
    batch 1, timed the way its own evaluator does (wall clock around compress and decompress, device synchronised)."""
    import torch
    from stf_b200.synth import synthetic_image
    net = _reference_net("cuda")
    times = []
    for i in range(warmup + steps):
        x = synthetic_image(1, H, W, seed=100 + i).cuda()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            enc = net.compress(x)
            net.decompress(enc["strings"], enc["shape"])
        torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return {"value": steps * H * W / sum(times) / 1e6, "unit": "Mpixel/s", "batch": 1,
            "what": "unmodified reference, model.to('cuda'), eager, 1 image 768x512 per step (context only)"}


def run_reference(args, rank):
    if rank != 0:
        return
    mpx, ms, base = cpu_reference_rate(args.steps, args.warmup)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": mpx, "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "STF compress+decompress, 1 image 768x512 per step on host CPU cores (bounded sample of the product arm's batch)",
                   "image": [H, W], "implementation": base["kind"]},
        "cpu_baseline": base,
        "e2e": {"value": mpx, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


# ----------------------------------------------------------------------------- product arm

def init_nccl_quiet(dist, dev):
    """NCCL process group + first collective with fd 1 pointed at stderr: whatever NCCL_DEBUG level the environment sets,
    NCCL's banner ("NCCL version ...", printed on stdout at communicator creation) stays out of the one-JSON-line stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        import torch
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from stf_b200 import ops
    from stf_b200.models import SymmetricalTransFormer
    from stf_b200.synth import synthetic_image

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        init_nccl_quiet(dist, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    net = SymmetricalTransFormer()
    torch.nn.Module.load_state_dict(net, synthetic_weights(), strict=False)
    net = net.to(dev).eval()
    net.update(force=True)
    B = args.batch
    # every step codes different images (seeded by rank and step): nothing can be cached across steps,
    # and one step touches >> L2 worth of activations (B * 98304 tokens * 48..384 ch fp32 per block)
    n_img = args.warmup + args.steps
    host_imgs = [synthetic_image(B, H, W, seed=1000 * rank + i).pin_memory() for i in range(n_img)]

    def step_device(x_dev):
        enc = net.compress(x_dev)
        dec = net.decompress(enc["strings"], enc["shape"])
        return enc, dec["x_hat"]

    out_host = torch.empty((B, 3, H, W), dtype=torch.float32, pin_memory=True)   # caller-owned, reused every step

    def step_e2e(x_host):
        x_dev = x_host.to(dev, non_blocking=True)       # H2D of this step's images (pinned source)
        enc, x_hat = step_device(x_dev)
        out_host.copy_(x_hat, non_blocking=True)         # D2H of the reconstructions
        torch.cuda.current_stream().synchronize()
        return enc, out_host

    def timed(fn, inputs):
        for i in range(args.warmup):
            fn(inputs[i])
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        l0 = ops.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        nbytes = 0
        for i in range(args.warmup, n_img):
            enc, _ = fn(inputs[i])
            nbytes += sum(len(s) for grp in enc["strings"] for s in grp)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        clocks = sampler.stop()
        launches = ops.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, clocks, nbytes

    dev_imgs = [x.to(dev) for x in host_imgs]
    torch.cuda.synchronize()
    ms_dev, launches, clocks, nbytes = timed(step_device, dev_imgs)
    ms_e2e, _, clocks_e2e, _ = timed(step_e2e, host_imgs)
    pixels = world * B * H * W * args.steps
    value = pixels / (ms_dev * 1e-3) / 1e6
    e2e = pixels / (ms_e2e * 1e-3) / 1e6

    # instrumented step (not timed for throughput): CUDA-event time per stf_b200 kernel family, every launch classed
    # HBM- or tensor-bound by its own arithmetic intensity (the GEMM engine serves both kinds of layer)
    roofline, by_kernel, tensor_pipe = None, None, None
    if rank == 0:
        from stf_b200 import profiler
        peaks = load_peaks()
        net.cuda_graphs = False          # eager launches so that every kernel can be bracketed by events
        step_device(dev_imgs[-1])
        with profiler.capture() as prof:
            step_device(dev_imgs[-1])
        torch.cuda.synchronize()
        net.cuda_graphs = True
        passes = 3 if ops.precision() == "fp32" else 1
        fam = prof.summary(ridge_flop_per_byte=peaks["tf32_tflops"] * 1e12 / passes / (peaks["hbm_gbs"] * 1e9))
        by_kernel = []
        for key, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
            per = {"kernel": f["name"], "bound": f["bound"], "launches_per_step": f["launches"], "ms_per_step": round(f["ms"], 3),
                   "avg_launch_us": f["ms"] * 1e3 / f["launches"]}
            if f["bound"] == "tensor":
                # algorithmic flops / time vs the TF32 tensor peak; the conv family runs single-pass TF32, the linear family
                # 3xTF32 in the fp32 mode (three MMAs per algorithmic product: pipe work = passes x algorithmic flops)
                p = 1 if f["name"].endswith(":conv") and ops.conv_precision_code() == 0 else passes
                ach = f["flops"] / (f["ms"] * 1e-3) / 1e12
                per.update({"achieved": ach, "peak": peaks["tf32_tflops"], "unit": "TFLOP/s", "mma_passes": p,
                            "frac": ach * p / peaks["tf32_tflops"], "algorithmic_flops_per_launch": f["flops"] / f["launches"]})
            else:
                ach = f["bytes"] / (f["ms"] * 1e-3) / 1e9
                per.update({"achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                            "algorithmic_bytes_per_launch": f["bytes"] / f["launches"]})
            per["traffic"], per["traffic_source"] = measured_traffic(f["name"], f["bound"], f["launches"], B)
            by_kernel.append(per)
        if by_kernel:
            roofline = dict(by_kernel[0])
            roofline.update({"peak_source": peaks["source"], "instrumented_step_ms": prof.total_ms,
                             "step_share": {f"{k['kernel']}|{k['bound']}": k["ms_per_step"] for k in by_kernel}})
            tens = [k for k in by_kernel if k["bound"] == "tensor"]
            if tens:   # the metric's second half: tensor-pipe utilisation of the window-attention GEMMs and convolutions
                t_ms = sum(k["ms_per_step"] for k in tens)
                tensor_pipe = {"util_estimate_pct": 100 * sum(k["frac"] * k["ms_per_step"] for k in tens) / t_ms,
                               "how": "time-weighted (algorithmic flops x MMA passes) / (CUDA-event time x TF32 peak) over the "
                                      "tensor-bound launches of one instrumented step; TF32 peak = measured bf16 peak / 2",
                               "ncu": ncu_tensor_pipe()}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, _, cpu_base = cpu_reference_rate(steps=3, warmup=1)

    extras = {}
    if not args.no_extras:
        del dev_imgs, host_imgs
        net._drop_plans()
        torch.cuda.empty_cache()
        for name, fn in (("train", train_record), ("wacnn", wacnn_record)):
            try:
                extras[name] = fn(rank, world, dev)
            except Exception as e:      # a sub-record never takes the headline line down
                extras[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            if world > 1:
                dist.barrier()
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            try:
                extras["reference_gpu_eager"] = gpu_eager_reference_rate()
            except Exception as e:
                extras["reference_gpu_eager"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (3xTF32 split on tcgen05, fp32 accumulate; convolutions single-pass TF32 like the "
                                          "reference's cuDNN default)" if ops.precision() == "fp32" else "tf32",
            "data": "synthetic",
            "config": {"workload": f"STF compress+decompress, batch {B} x 768x512 RGB per GPU per step (BASELINE config 3)",
                       "batch_per_gpu": B, "image": [H, W], "gemm_precision": ops.precision(),
                       "conv_precision": "fp32" if ops.conv_precision_code() else "tf32",
                       "weights": "synthetic (stf_b200/synth.py seed 0)",
                       "l2": "inputs differ every step; per-step activations >> 126 MB L2",
                       "bpp": nbytes * 8 / (B * H * W * args.steps)},
            "e2e": {"value": e2e, "unit": "Mpixel/s", "h2d_bytes_per_step": B * 3 * H * W * 4,
                    "d2h_bytes_per_step": B * 3 * H * W * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "clocks_e2e": clocks_e2e,
            "roofline": roofline, "roofline_by_kernel": by_kernel, "tensor_pipe": tensor_pipe, "cpu_baseline": cpu_base,
            "extras": extras}), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- sub-records (BASELINE configs 5 and 4)

def train_record(rank, world, dev, batch=16, steps=5, warmup=3):
    """BASELINE config 5: STF rate-distortion training step (lambda 0.0035, 16 x 256x256 per GPU, Adam + aux Adam, clip 1.0),
    NCCL gradient all-reduce over NVLink at N > 1.  images/s over all ranks (weak scaling), CUDA events, max over ranks;
    exposed communication = step time minus the same step with the all-reduce switched off (gradients left unsynchronised,
    timing only)."""
    import torch
    import torch.distributed as dist
    from stf_b200 import ops
    from stf_b200.models import SymmetricalTransFormer
    from stf_b200.synth import synthetic_image
    from stf_b200.training import GradientAllReduce, RateDistortionLoss, configure_optimizers, train_step
    torch.manual_seed(0)                                   # identical replicas on every rank
    net = SymmetricalTransFormer()
    torch.nn.Module.load_state_dict(net, synthetic_weights(), strict=False)
    net = net.to(dev).train()
    opt, aux = configure_optimizers(net, 1e-4, 1e-3)
    crit = RateDistortionLoss(0.0035)
    red = GradientAllReduce(net.parameters()).attach() if world > 1 else None
    n = warmup + steps
    imgs = [synthetic_image(batch, 256, 256, seed=1000 * rank + i).to(dev) for i in range(n)]
    torch.manual_seed(100 + rank)                           # per-rank noise / stochastic-depth streams

    def run(reducer, k0, k1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for i in range(k0, k1):
            out = train_step(net, imgs[i % n], crit, opt, aux, reducer)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    run(red, 0, warmup)
    l0 = ops.launch_count()
    ms, out = run(red, warmup, n)
    launches = ops.launch_count() - l0
    rec = {"metric": "STF rate-distortion training step, images/s (BASELINE config 5)", "value": world * batch * steps / (ms * 1e-3),
           "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "scaling": "weak",
           "config": {"workload": f"batch {batch} x 256x256 per GPU, lambda 0.0035, Adam 1e-4 + aux Adam 1e-3, clip 1.0",
                      "gemm_precision": ops.precision()},
           "loss": float(out["loss"].detach()), "gpu_launches": launches}
    if world > 1:
        n_params = sum(p.numel() for p in net.parameters())
        ms_nc, _ = run(None, 0, steps)          # same step without the collective (timing only; replicas diverge afterwards)
        rec.update({"allreduce_bytes_per_step": 4 * n_params, "collective": "NCCL all-reduce (mean) of all gradients, 50 MB buckets "
                    "launched from backward hooks", "ms_per_step_without_allreduce": ms_nc / steps,
                    "exposed_comm_ms": max(0.0, (ms - ms_nc) / steps)})
    return rec


def wacnn_record(rank, world, dev, batch=1, steps=2, warmup=2):
    """BASELINE config 4: WACNN compress + decompress at 2048x1408, `batch` images per GPU, batch-sharded over the ranks (no
    data-path collective).  Mpixel/s over all ranks, CUDA events, max over ranks."""
    import torch
    import torch.distributed as dist
    from stf_b200 import ops
    from stf_b200.models import WACNN
    from stf_b200.synth import synthetic_image, synthetic_state_dict
    Hc, Wc = 1408, 2048
    spec = {k: (tuple(s), getattr(torch, d.split(".")[-1])) for k, (s, d) in
            json.load(open(os.path.join(ROOT, "tests", "golden", "cnn_spec.json"))).items()}
    net = WACNN()
    torch.nn.Module.load_state_dict(net, synthetic_state_dict(spec, 0), strict=False)
    net = net.to(dev).eval()
    net.update(force=True)
    n = warmup + steps
    imgs = [synthetic_image(batch, Hc, Wc, seed=5000 + 100 * rank + i).to(dev) for i in range(n)]

    def step(x):
        enc = net.compress(x)
        net.decompress(enc["strings"], enc["shape"])
        return enc

    for i in range(warmup):
        step(imgs[i])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nbytes = 0
    for i in range(warmup, n):
        enc = step(imgs[i])
        nbytes += sum(len(s) for g in enc["strings"] for s in g)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    px = world * batch * Hc * Wc * steps
    return {"metric": "WACNN encode+decode Mpixel/s at 2048x1408 (BASELINE config 4)", "value": px / (ms * 1e-3) / 1e6,
            "unit": "Mpixel/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "scaling": "weak",
            "config": {"workload": f"WACNN compress+decompress, batch {batch} x 2048x1408 per GPU", "bpp": nbytes * 8 / (batch * Hc * Wc * steps)},
            "gpu_launches": ops.launch_count() - l0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config 5 (training) and config 4 (WACNN) sub-records")
    ap.add_argument("--precision", default=None, choices=["fp32", "tf32"],
                    help="GEMM arithmetic: fp32 = 3xTF32 split (default, parity with the reference's fp32 matmuls), tf32 = single pass")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # convenience: `python bench.py --gpus N` re-launches itself under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"), __file__] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    args.warmup = max(args.warmup, 3)
    if args.precision:
        os.environ["STF_B200_PRECISION"] = args.precision
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
