"""Synchronised wall-clock breakdown of STF compress / decompress (developer tool).
   python tools/phase_breakdown.py [--batch 8] [--iters 3]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from stf_b200 import models, ops  # noqa: E402
from stf_b200.synth import synthetic_image  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=3)
    args = ap.parse_args()
    net = models.SymmetricalTransFormer()
    torch.nn.Module.load_state_dict(net, bench.synthetic_weights(), strict=False)
    net = net.cuda().eval()
    net.update(force=True)
    xs = [synthetic_image(args.batch, bench.H, bench.W, seed=i).cuda() for i in range(args.iters + 1)]
    enc = net.compress(xs[0])
    net.decompress(enc["strings"], enc["shape"])
    models.PHASE_TIMES = {}
    l0 = ops.launch_count()
    for x in xs[1:]:
        enc = net.compress(x)
        net.decompress(enc["strings"], enc["shape"])
    tot = 0.0
    for k, v in models.PHASE_TIMES.items():
        print(f"{k:20s} {v / args.iters:8.2f} ms")
        tot += v / args.iters
    print(f"{'sum':20s} {tot:8.2f} ms per step of {args.batch} images; stf_b200 launches/step {(ops.launch_count() - l0) / args.iters:.0f}")
    print("host threads:", os.cpu_count(), " y bytes/image:", sum(len(s) for s in enc["strings"][0]) / args.batch)


if __name__ == "__main__":
    main()
