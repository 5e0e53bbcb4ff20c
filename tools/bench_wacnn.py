#!/usr/bin/env python
"""BASELINE config 4: WACNN (CNN model with WinBasedAttention) compress + decompress at CLIC size 2048x1408, batch per GPU.
   python tools/bench_wacnn.py [--batch 2] [--steps 3] [--warmup 2]
One JSON line: Mpixel/s with inputs in HBM (`value`) and through host buffers (`e2e`), plus the share of the stf_b200 kernels
(window attention with 64-token / head_dim-24 and 16-token / head_dim-40 windows, entropy kernels) in one instrumented step."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from stf_b200 import ops, profiler  # noqa: E402
from stf_b200.models import WACNN  # noqa: E402
from stf_b200.synth import synthetic_image, synthetic_state_dict  # noqa: E402

H, W = 1408, 2048


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    spec = {k: (tuple(s), getattr(torch, d.split(".")[-1])) for k, (s, d) in
            json.load(open(os.path.join(ROOT, "tests", "golden", "cnn_spec.json"))).items()}
    net = WACNN()
    torch.nn.Module.load_state_dict(net, synthetic_state_dict(spec, 0), strict=False)
    net = net.cuda().eval()
    net.update(force=True)
    n = args.warmup + args.steps
    host = [synthetic_image(args.batch, H, W, seed=i).pin_memory() for i in range(n)]
    dev = [x.cuda() for x in host]
    out_host = torch.empty((args.batch, 3, H, W), pin_memory=True)

    def step(x):
        enc = net.compress(x)
        return enc, net.decompress(enc["strings"], enc["shape"])["x_hat"]

    def step_e2e(xh):
        enc, xhat = step(xh.cuda(non_blocking=True))
        out_host.copy_(xhat, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return enc

    def timed(fn, xs):
        for x in xs[:args.warmup]:
            fn(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.launch_count()
        e0.record()
        nbytes = 0
        for x in xs[args.warmup:]:
            r = fn(x)
            enc = r[0] if isinstance(r, tuple) else r
            nbytes += sum(len(s) for g in enc["strings"] for s in g)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), ops.launch_count() - l0, nbytes

    ms, launches, nbytes = timed(step, dev)
    ms_e2e, _, _ = timed(step_e2e, host)
    px = args.batch * H * W * args.steps
    net.cuda_graphs = False
    step(dev[-1])
    with profiler.capture() as prof:
        step(dev[-1])
    torch.cuda.synchronize()
    fam = prof.summary()
    print(json.dumps({
        "metric": "WACNN encode+decode Mpixel/s at 2048x1408", "value": px / (ms * 1e-3) / 1e6, "unit": "Mpixel/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "dtype": ops.precision(), "data": "synthetic",
        "config": {"workload": f"WACNN compress+decompress, batch {args.batch} x 2048x1408 (BASELINE config 4)",
                   "bpp": nbytes * 8 / px},
        "e2e": {"value": px / (ms_e2e * 1e-3) / 1e6, "unit": "Mpixel/s", "h2d_bytes_per_step": args.batch * 3 * H * W * 4,
                "d2h_bytes_per_step": args.batch * 3 * H * W * 4},
        "gpu_launches": launches,
        "stf_b200_kernel_ms_in_one_eager_step": {k: round(v["ms"], 3) for k, v in fam.items()},
        "eager_step_ms": prof.total_ms}))


if __name__ == "__main__":
    main()
