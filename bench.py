#!/usr/bin/env python
"""bench.py -- STF encode+decode throughput at 768x512 (BASELINE.json metric, config 3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A "step" = compress() + decompress() of one batch of B synthetic 768x512 RGB images per GPU (weak
scaling: every rank codes its own B images, no data-path collective).
One JSON line is printed by rank 0:
  value        Mpixel/s, whole job, inputs resident in HBM at the start of the timed region
  e2e          same through the public API with HOST buffers: pinned-host images -> H2D -> compress ->
               byte strings -> decompress -> x_hat -> D2H, all inside the timed region
  roofline     dominant stf_b200 kernel (by device time over one instrumented step): algorithmic bytes
               per launch / CUDA-event time per launch vs the measured HBM peak
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
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(kernel, batch):
    """DRAM bytes per launch of `kernel` (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu
    capture of one step (profiles/r1_traffic_linear_step.json, batch 8), scaled linearly to this batch."""
    p = os.path.join(ROOT, "profiles", "r1_traffic_linear_step.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    if kernel not in d:
        return None, None
    k = d[kernel]
    per_launch = (k["dram_read_bytes"] + k["dram_write_bytes"]) / k["launches"] * batch / k["batch"]
    return per_launch, f"ncu capture at batch {k['batch']} scaled x{batch / k['batch']:g} (profiles/r1_traffic_linear_step.json)"


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

    # instrumented step (not timed for throughput): CUDA-event time per stf_b200 kernel family
    roofline = None
    if rank == 0:
        from stf_b200 import profiler
        peak, peak_src = load_peaks()
        net.cuda_graphs = False          # eager launches so that every kernel can be bracketed by events
        step_device(dev_imgs[-1])
        with profiler.capture() as prof:
            step_device(dev_imgs[-1])
        torch.cuda.synchronize()
        net.cuda_graphs = True
        fam = prof.summary()
        if fam:
            top = max(fam.values(), key=lambda f: f["ms"])
            achieved = top["bytes"] / (top["ms"] * 1e-3) / 1e9
            traffic, traffic_src = measured_traffic(top["name"], B)
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "traffic_source": traffic_src, "kernel": top["name"],
                        "launches_per_step": top["launches"],
                        "avg_launch_us": top["ms"] * 1e3 / top["launches"],
                        "algorithmic_bytes_per_launch": top["bytes"] / top["launches"], "peak_source": peak_src,
                        "step_share": {k: round(v["ms"], 3) for k, v in fam.items()},
                        "instrumented_step_ms": prof.total_ms}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, _, cpu_base = cpu_reference_rate(steps=3, warmup=1)

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (3xTF32 split on tcgen05, fp32 accumulate)" if ops.precision() == "fp32" else "tf32",
            "data": "synthetic",
            "config": {"workload": f"STF compress+decompress, batch {B} x 768x512 RGB per GPU per step (BASELINE config 3)",
                       "batch_per_gpu": B, "image": [H, W], "gemm_precision": ops.precision(), "weights": "synthetic (stf_b200/synth.py seed 0)",
                       "l2": "inputs differ every step; per-step activations >> 126 MB L2",
                       "bpp": nbytes * 8 / (B * H * W * args.steps)},
            "e2e": {"value": e2e, "unit": "Mpixel/s", "h2d_bytes_per_step": B * 3 * H * W * 4,
                    "d2h_bytes_per_step": B * 3 * H * W * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "clocks_e2e": clocks_e2e,
            "roofline": roofline, "cpu_baseline": cpu_base}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
