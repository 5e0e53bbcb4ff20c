#!/usr/bin/env python
"""BASELINE configs 1 and 2: STF eval-mode forward (x_hat + y / z likelihoods).
   config 2: batch 16 x 256x256 on one B200 through the stf_b200 kernels (CUDA events, inputs in HBM, L2 flushed by the
             >126 MB of activations per pass); config 1: the CPU oracle (the reference's own CPU-runnable case), 1 x 256x256.
   python tools/bench_forward.py [--batch 16] [--steps 10] [--warmup 3]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from stf_b200 import ops  # noqa: E402
from stf_b200.models import SymmetricalTransFormer  # noqa: E402
from stf_b200.synth import synthetic_image  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    sd = bench.synthetic_weights()
    net = SymmetricalTransFormer()
    torch.nn.Module.load_state_dict(net, sd, strict=False)
    net = net.cuda().eval()
    xs = [synthetic_image(args.batch, 256, 256, seed=i).cuda() for i in range(args.warmup + args.steps)]
    for x in xs[:args.warmup]:
        out = net(x)
    torch.cuda.synchronize()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for x in xs[args.warmup:]:
        out = net(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    bpp = float((-torch.log2(out["likelihoods"]["y"]).sum() - torch.log2(out["likelihoods"]["z"]).sum()) / (args.batch * 65536))
    # config 1: CPU oracle, one 256x256 image
    from oracle import codec as OC
    torch.set_num_threads(os.cpu_count() or 1)
    ora = OC.StfOracle(sd)
    x1 = synthetic_image(1, 256, 256, seed=0)
    ora.forward(x1)
    t0 = time.perf_counter()
    for _ in range(3):
        ora.forward(x1)
    cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    print(json.dumps({
        "metric": "STF forward + likelihoods, images/s (config 2)", "value": args.batch / (ms * 1e-3), "unit": "images/s",
        "ms_per_step": ms, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "dtype": ops.precision(), "data": "synthetic",
        "config": {"workload": f"STF eval forward, batch {args.batch} x 256x256 (BASELINE config 2)", "bpp_estimate": bpp},
        "gpu_launches": ops.launch_count() - l0,
        "cpu_baseline": {"value": 1e3 / cpu_ms, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "3 x forward of 1 x 256x256 (BASELINE config 1: oracle port on torch CPU ops)", "ms_per_image": cpu_ms}}))


if __name__ == "__main__":
    main()
