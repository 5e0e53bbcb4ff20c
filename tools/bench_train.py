#!/usr/bin/env python
"""Training-step throughput (BASELINE config 5): STF rate-distortion step, lambda 0.0035, 16 x 256x256 per GPU,
Adam + aux Adam, gradient clipping 1.0, NCCL gradient all-reduce over NVLink at N > 1.

    python tools/bench_train.py [--steps K] [--warmup W] [--batch 16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py --gpus N

One JSON line from rank 0: images/s over all ranks (weak scaling: per-GPU batch fixed), ms per step (CUDA events,
max over ranks), all-reduced bytes per step, and -- at N = 1 -- the CPU oracle's training step (torch CPU autograd over
the restated reference, all host threads) on a bounded sample for scale.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cpu_step_rate(batch, steps=1):
    import math

    import torch

    import bench
    from oracle import codec as OC
    from stf_b200.synth import synthetic_image
    torch.set_num_threads(os.cpu_count() or 1)
    ora = OC.StfOracle(bench.synthetic_weights())
    for v in ora.sd.values():
        if v.is_floating_point() and v.dim() > 0:
            v.requires_grad_(True)
    x = synthetic_image(batch, 256, 256, seed=1)
    noise = {"y": torch.rand(batch, 384, 16, 16) - 0.5, "z": torch.rand(batch, 192, 4, 4) - 0.5}
    t0 = time.perf_counter()
    for _ in range(steps):
        out = ora.forward_train(x, noise)
        bpp = sum(torch.log(l).sum() / (-math.log(2) * batch * 256 * 256) for l in out["likelihoods"].values())
        (0.0035 * 255 ** 2 * torch.nn.functional.mse_loss(out["x_hat"], x) + bpp).backward()
    dt = (time.perf_counter() - t0) / steps
    return {"value": batch / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} x forward+backward of {batch} x 256x256 (oracle restatement on torch CPU autograd, no optimizer)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reducer", default="buckets", choices=["ddp", "buckets"],
                    help="N > 1: torch DistributedDataParallel (train.py:363) or stf_b200.training.GradientAllReduce")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import bench
    from stf_b200 import ops
    from stf_b200.models import SymmetricalTransFormer
    from stf_b200.synth import synthetic_image
    from stf_b200.training import GradientAllReduce, RateDistortionLoss, configure_optimizers, train_step
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        bench.init_nccl_quiet(dist, dev)
    torch.manual_seed(0)                                   # identical replicas on every rank
    net = SymmetricalTransFormer()                          # constructor defaults: drop_path_rate 0.2 live in train()
    torch.nn.Module.load_state_dict(net, bench.synthetic_weights(), strict=False)
    net = net.to(dev).train()
    opt, aux = configure_optimizers(net, 1e-4, 1e-3)
    crit = RateDistortionLoss(0.0035)
    red, model = None, net
    if world > 1 and args.reducer == "buckets":
        red = GradientAllReduce(net.parameters()).attach()
    elif world > 1:
        net._prepare_inference()     # conv weights to channels_last BEFORE DDP fixes its bucket-view strides
        model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], bucket_cap_mb=50,
                                                          gradient_as_bucket_view=True)
    n = args.warmup + args.steps
    imgs = [synthetic_image(args.batch, 256, 256, seed=1000 * rank + i).to(dev) for i in range(n)]
    torch.manual_seed(100 + rank)                           # per-rank noise / stochastic-depth streams

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        out = train_step(model, imgs[i], crit, opt, aux, red)
    barrier()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.warmup, n):
        out = train_step(model, imgs[i], crit, opt, aux, red)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        n_params = sum(p.numel() for p in net.parameters())
        line = {"metric": "STF rate-distortion training step, images/s (config 5)", "value": world * args.batch * args.steps / (ms * 1e-3),
                "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "dtype": "f32 (3xTF32 GEMMs)" if ops.precision() == "fp32" else "tf32", "data": "synthetic",
                "config": {"workload": f"STF train step, batch {args.batch} x 256x256 per GPU, lambda 0.0035, Adam 1e-4 + aux Adam 1e-3, clip 1.0",
                           "parameters": n_params, "allreduce_bytes_per_step": 4 * n_params if world > 1 else 0,
                           "collective": ("NCCL all-reduce (mean) of all gradients, 50 MB buckets, " +
                                          ("torch DDP reducer (overlapped with backward)" if args.reducer == "ddp" else
                                           "GradientAllReduce hooks")) if world > 1 else "none"},
                "loss": float(out["loss"].detach()), "gpu_launches": ops.launch_count() - l0,
                "peak_memory_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_step_rate(2)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
